mkdir -p gpurun_out
for v in default pb2o8 pb4o5 pb8o4; do
  if [ $v = default ]; then unset H2AES_LIB; else export H2AES_LIB=$PWD/halo2_aes_b200/lib/variants/libh2aes_$v.so; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/ab_$v.json 2> gpurun_out/ab.err
  python -c "
import json; d=json.load(open('gpurun_out/ab_$v.json')); print('$v', round(d['value']), round(d['ms_per_step'],3), {k: round(x,3) for k,x in d['phase_ms'].items()})"
done
unset H2AES_LIB
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged or sharded or key_schedule" 2>&1 | tail -2
