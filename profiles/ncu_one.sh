# one `ncu --set full` capture of a kernel of the bench step: bash profiles/ncu_one.sh REGEX NAME [skip]
# (run only after the same bench command exited 0 without ncu)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$1" -s ${3:-8} -c ${4:-1} \
    -f -o gpurun_out/$2 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-prover-side --no-perm \
    > gpurun_out/ncu_$2.log 2>&1; echo ncu=$?
ncu -i gpurun_out/$2.ncu-rep --page raw --csv > gpurun_out/$2_raw.csv 2>/dev/null
ncu -i gpurun_out/$2.ncu-rep --page source --csv > gpurun_out/$2_source.csv 2>/dev/null
