# Round-2 evidence on one B200: bench line, reference arm, ncu launch list of the hot path and
# `ncu --set full` captures of its top kernels, summarised on the box.  Run under gpurun from the
# repository root (only after the same commands exited 0 without ncu); results land in
# gpurun_out/ and are copied to profiles/ afterwards.
mkdir -p gpurun_out
T=${TAG:-r2_final}
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo bench=$?
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_reference_arm.json 2> gpurun_out/${T}_ref.err; echo ref=$?
B="python bench.py --steps 2 --warmup 3 --circuits 2 --no-cpu --no-e2e --no-prover-side --no-perm --no-config4 --no-sharded"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_ncu.csv $B > gpurun_out/ncu1.log 2>&1; echo launches=$?
python profiles/summarize_launches.py gpurun_out/${T}_launches_ncu.csv > gpurun_out/${T}_launches_summary.txt
K="k_advice|k_permuted|k_general_list|k_inv_tables|k_z_general|k_z_totals|k_z_prefix|k_z_write|k_lookup_hist|k_bucket_rank"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$K" --launch-skip 40 -c 10 -f -o /tmp/${T}_prof $B > gpurun_out/ncu3.log 2>&1; echo full=$?
for k in k_permuted k_z_write k_z_totals k_z_general k_z_prefix k_advice; do python profiles/ncu_summary.py /tmp/${T}_prof.ncu-rep $k > /tmp/s_$k.txt 2>&1; done
( sed -n '1,/==== source page/p' /tmp/s_k_permuted.txt | grep -v "==== source page"; for k in k_permuted k_z_write k_z_totals k_z_general k_z_prefix k_advice; do sed -n '/==== source page/,$p' /tmp/s_$k.txt; done ) > gpurun_out/${T}_ncu_full_summary.txt
python profiles/make_traffic.py /tmp/${T}_prof.ncu-rep "profiles/${T}_ncu_full_summary.txt (K=20,N=5, 3853 blocks, one launch each)" > gpurun_out/${T}_ncu_traffic.json
tail -n 2 gpurun_out/ncu3.log; du -sh gpurun_out
