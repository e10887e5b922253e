//! `src/b200/mod.rs` of tkmct/halo2-aes with `feature = "b200"`: the host side of the drop-in.
//!
//! UNCOMPILED SKETCH — this environment has no Rust toolchain and halo2_proofs is not vendored in
//! the reference (Cargo.toml:14-18), so nothing here has been through rustc.  It is the shim
//! INTEGRATION.md describes, written out: `FixedAes128Config::{schedule_key, encrypt}` record
//! their inputs here instead of assigning cells (src/aes128.rs:143-265), one `flush` per
//! `synthesize` asks the GPU for every advice column, and `install` hands the cells and the copy
//! constraints to halo2 so that MockProver / keygen / create_proof run unmodified.
//!
//! `ffi.rs` (generated from include/h2aes.h) declares the C ABI; `build.rs` compiles the CUDA
//! sources.

pub mod ffi;

use std::collections::HashMap;
use std::ptr;

use halo2_proofs::circuit::{AssignedCell, Cell, Layouter, Region, Value};
use halo2_proofs::plonk::{Advice, Column, Error};
use halo2curves::bn256::Fr as Fp;

use ffi::*;

/// What `schedule_key` / `encrypt` leave behind until `flush`.
pub struct B200State {
    ctx: *mut h2aes_ctx,
    circuit: *mut h2aes_circuit,
    pub layout: h2aes_layout,
    pub key: Option<[u8; 16]>,
    pub plaintexts: Vec<[u8; 16]>,
}

fn check(rc: i32) -> Result<(), Error> {
    match rc {
        H2AES_OK => Ok(()),
        // src/aes128.rs:160-162
        H2AES_ERR_CAPACITY => panic!("AES calls too many. doesn't fit in the rows"),
        // src/aes128.rs:170
        H2AES_ERR_NOT_SCHEDULED => panic!("Keys should be scheduled"),
        H2AES_ERR_ROWS => Err(Error::NotEnoughRowsAvailable { current_k: 0 }),
        H2AES_ERR_LOOKUP => Err(Error::ConstraintSystemFailure),
        _ => Err(Error::Synthesis),
    }
}

impl B200State {
    /// `K`, `N` as in `FixedAes128Config<K, N>`; `device` = CUDA ordinal.
    pub fn new(k: u32, n: usize, device: i32) -> Result<Self, Error> {
        let mut layout = h2aes_layout::default();
        let (mut ctx, mut circuit) = (ptr::null_mut(), ptr::null_mut());
        unsafe {
            check(h2aes_layout_query(H2AES_KIND_AES, k as i32, n as i32, &mut layout))?;
            check(h2aes_ctx_create(device, H2AES_SBOX_REFERENCE, &mut ctx))?;
            let rc = h2aes_circuit_create(ctx, H2AES_KIND_AES, k as i32, n as i32, &mut circuit);
            if rc != H2AES_OK {
                h2aes_ctx_destroy(ctx);
                check(rc)?;
            }
        }
        Ok(Self { ctx, circuit, layout, key: None, plaintexts: Vec::new() })
    }

    /// replaces the body of `schedule_key` (src/aes128.rs:143-152)
    pub fn schedule_key(&mut self, key: [u8; 16]) {
        self.key = Some(key);
    }

    /// replaces the body of `encrypt` (src/aes128.rs:154-265): capacity rule first (`aes_callable`,
    /// :303-325), then the block is only recorded.  Returns (column set, first row) of the block.
    pub fn encrypt(&mut self, plaintext: [u8; 16]) -> (usize, usize) {
        self.key.expect("Keys should be scheduled");
        let (mut set, mut row) = (0i32, 0u64);
        let rc = unsafe {
            h2aes_block_position(self.layout.k, self.layout.n_sets, self.plaintexts.len() as u64,
                                 &mut set, &mut row)
        };
        if rc == H2AES_ERR_CAPACITY {
            panic!("AES calls too many. doesn't fit in the rows");
        }
        self.plaintexts.push(plaintext);
        (set as usize, row as usize)
    }

    /// One GPU call for the whole circuit: every advice column, column-major, and the ciphertexts.
    pub fn flush(&mut self) -> Result<(Vec<Fp>, Vec<[u8; 16]>), Error> {
        let key = self.key.expect("Keys should be scheduled");
        let cells = self.layout.n_advice as usize * self.layout.n_rows as usize;
        let mut advice = vec![Fp::zero(); cells];
        let mut cts = vec![[0u8; 16]; self.plaintexts.len()];
        check(unsafe {
            h2aes_witness_host(self.circuit, key.as_ptr(), self.plaintexts.as_ptr() as *const u8,
                               self.plaintexts.len() as u64, advice.as_mut_ptr() as *mut u64,
                               cts.as_mut_ptr() as *mut u8)
        })?;
        Ok((advice, cts))
    }

    /// Hands the columns to stock halo2: one region over all rows, `assign_advice` for every cell
    /// the chips would have assigned (non-zero values and both ends of every copy), then
    /// `constrain_equal` per copy pair.  A prover that accepts whole columns skips this.
    pub fn install(&self, layouter: &mut impl Layouter<Fp>, columns: &[Column<Advice>], advice: &[Fp])
                   -> Result<HashMap<(usize, usize), AssignedCell<Fp, Fp>>, Error> {
        let n = self.layout.n_rows as usize;
        let mut n_copies = 0u64;
        check(unsafe { h2aes_copy_count(H2AES_KIND_AES, self.plaintexts.len() as u64, &mut n_copies) })?;
        let mut copies = vec![[0u64; 4]; n_copies as usize];
        check(unsafe {
            h2aes_copy_constraints_host(H2AES_KIND_AES, self.layout.k, self.layout.n_sets,
                                        self.plaintexts.len() as u64, 0, n_copies,
                                        copies.as_mut_ptr() as *mut u64)
        })?;
        layouter.assign_region(
            || "b200 witness",
            |mut region: Region<'_, Fp>| {
                let mut assigned: HashMap<(usize, usize), AssignedCell<Fp, Fp>> = HashMap::new();
                let mut put = |region: &mut Region<'_, Fp>, col: usize, row: usize|
                               -> Result<Cell, Error> {
                    if let Some(c) = assigned.get(&(col, row)) {
                        return Ok(c.cell());
                    }
                    let v = advice[col * n + row];
                    let c = region.assign_advice(|| "", columns[col], row, || Value::known(v))?;
                    let cell = c.cell();
                    assigned.insert((col, row), c);
                    Ok(cell)
                };
                for col in 0..columns.len() {
                    for row in 0..self.layout.usable_rows as usize {
                        if advice[col * n + row] != Fp::zero() {
                            put(&mut region, col, row)?;
                        }
                    }
                }
                for c in &copies {
                    let a = put(&mut region, c[0] as usize, c[1] as usize)?;
                    let b = put(&mut region, c[2] as usize, c[3] as usize)?;
                    region.constrain_equal(a, b)?;
                }
                Ok(assigned)
            },
        )
    }

    /// The 16 ciphertext cells `encrypt` returns for block `b`: a2 of rows 1344..1359 of the block.
    pub fn ciphertext_cells(&self, assigned: &HashMap<(usize, usize), AssignedCell<Fp, Fp>>, b: usize)
                            -> Vec<AssignedCell<Fp, Fp>> {
        let (mut set, mut row) = (0i32, 0u64);
        unsafe { h2aes_block_position(self.layout.k, self.layout.n_sets, b as u64, &mut set, &mut row) };
        (0..16).map(|i| assigned[&(3 * set as usize + 2, row as usize + 1344 + i)].clone()).collect()
    }

    /// The two lookup phases of `create_proof` (plonk/lookup/prover.rs): after theta, after beta/gamma.
    pub fn lookup_permuted(&mut self, theta: &Fp) -> Result<(Vec<Fp>, Vec<Fp>), Error> {
        let cells = self.layout.n_lookups as usize * self.layout.n_rows as usize;
        let (mut ap, mut sp) = (vec![Fp::zero(); cells], vec![Fp::zero(); cells]);
        check(unsafe {
            h2aes_lookup_permuted_host(self.circuit, theta as *const Fp as *const u64, 0,
                                       self.layout.n_lookups, ptr::null_mut(), ptr::null_mut(),
                                       ap.as_mut_ptr() as *mut u64, sp.as_mut_ptr() as *mut u64)
        })?;
        Ok((ap, sp))
    }
    pub fn lookup_product(&mut self, beta: &Fp, gamma: &Fp) -> Result<Vec<Fp>, Error> {
        let cells = self.layout.n_lookups as usize * self.layout.n_rows as usize;
        let mut z = vec![Fp::zero(); cells];
        check(unsafe {
            h2aes_lookup_product_host(self.circuit, beta as *const Fp as *const u64,
                                      gamma as *const Fp as *const u64, 0, self.layout.n_lookups,
                                      z.as_mut_ptr() as *mut u64)
        })?;
        Ok(z)
    }
}

impl Drop for B200State {
    fn drop(&mut self) {
        unsafe {
            h2aes_circuit_destroy(self.circuit);
            h2aes_ctx_destroy(self.ctx);
        }
    }
}
