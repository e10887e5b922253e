# Round-end evidence on one B200: bench line, reference arm, ncu launch lists (hot path and
# prover-side steps) and `ncu --set full` captures of the top kernels, summarised on the box
# (the reports themselves stay in /tmp unless small).  Run under gpurun from the repository
# root; results land in gpurun_out/ and are copied to profiles/ afterwards.
mkdir -p gpurun_out
T=${TAG:-r1_final}
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo bench=$?
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_reference_arm.json 2> gpurun_out/${T}_ref.err; echo ref=$?
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-prover-side"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_ncu.csv $B > gpurun_out/ncu1.log 2>&1; echo launches=$?
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file /tmp/ps_launches.csv python profiles/prover_side.py > gpurun_out/ncu2.log 2>&1; echo launches_ps=$?
python profiles/summarize_launches.py gpurun_out/${T}_launches_ncu.csv > gpurun_out/${T}_launches_summary.txt
python profiles/summarize_launches.py /tmp/ps_launches.csv > gpurun_out/${T}_prover_side_launches_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_advice|k_permuted|k_inv_tables|k_z_fill|k_z_chunks|k_z_propagate|k_z_down|k_lookup_hist|k_bucket_rank" --launch-skip 18 -c 9 -f -o /tmp/${T}_prof $B --no-perm > gpurun_out/ncu3.log 2>&1; echo full=$?
timeout 900 ncu --set full --clock-control none -k "regex:k_ntt_cols|k_ntt_rows|k_lookup_quotient|k_gate_perm_quotient|k_msm_tasks|k_msm_chunks" --launch-skip 40 -c 14 -f -o /tmp/${T}_prof_ps python profiles/prover_side.py > gpurun_out/ncu4.log 2>&1; echo full_ps=$?
for k in k_permuted k_z_fill k_z_propagate k_z_chunks k_advice; do python profiles/ncu_summary.py /tmp/${T}_prof.ncu-rep $k > /tmp/s_$k.txt 2>&1; done
( sed -n '1,/==== source page/p' /tmp/s_k_permuted.txt | grep -v "==== source page"; for k in k_permuted k_z_fill k_z_propagate k_z_chunks k_advice; do sed -n '/==== source page/,$p' /tmp/s_$k.txt; done ) > gpurun_out/${T}_ncu_full_summary.txt
python profiles/ncu_summary.py /tmp/${T}_prof_ps.ncu-rep k_ntt_rows | sed -n '1,/==== source page/p' | grep -v "==== source page" > gpurun_out/${T}_prover_side_ncu_full_summary.txt
python profiles/make_traffic.py /tmp/${T}_prof.ncu-rep "profiles/${T}_ncu_full_summary.txt (K=20,N=5, 3853 blocks, one launch each)" > gpurun_out/${T}_ncu_traffic.json
python profiles/make_traffic.py /tmp/${T}_prof_ps.ncu-rep "profiles/${T}_prover_side_ncu_full_summary.txt" > gpurun_out/${T}_prover_side_ncu_traffic.json
ls -la /tmp/*.ncu-rep
sz=$(stat -c %s /tmp/${T}_prof.ncu-rep); if [ "$sz" -lt 30000000 ]; then cp /tmp/${T}_prof.ncu-rep gpurun_out/; fi
tail -n 2 gpurun_out/ncu3.log; tail -n 2 gpurun_out/ncu4.log; du -sh gpurun_out
