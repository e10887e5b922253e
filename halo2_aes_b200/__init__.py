"""halo2_aes_b200 — B200-native witness-generation backend for tkmct/halo2-aes.

Produces on the GPU, bit-exactly, the bn256 Fr columns that the reference's AES-128 /
key-schedule chips assign on one CPU thread (src/aes128.rs, src/key_schedule.rs, src/chips/*,
src/table.rs) plus the halo2 lookup-argument helper columns (permuted input/table, grand
product Z).  The product is the C-ABI library `lib/libh2aes.so` (include/h2aes.h, sources in
csrc/); this package is the host-side mirror of the reference interface on top of it.
"""
from ._ffi import (CapacityError, H2AesError, KeysNotScheduled, NotEnoughRows, KIND_AES,
                   KIND_KEYSCHED, SBOX_FIPS197, SBOX_REFERENCE, block_position, layout_query)

__all__ = [
    "CapacityError", "H2AesError", "KeysNotScheduled", "NotEnoughRows", "KIND_AES",
    "KIND_KEYSCHED", "SBOX_FIPS197", "SBOX_REFERENCE", "block_position", "layout_query",
]
