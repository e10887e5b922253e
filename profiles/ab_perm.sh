for v in "$@"; do
  if [ $v = default ]; then unset H2AES_LIB; else export H2AES_LIB=$PWD/halo2_aes_b200/lib/variants/libh2aes_$v.so; fi
  timeout 600 python bench.py --no-cpu --no-e2e --steps 50 > gpurun_out/abp_$v.json 2> gpurun_out/abp.err || tail -3 gpurun_out/abp.err
  python -c "
import json; d=json.load(open('gpurun_out/abp_$v.json')); p=d['permutation_product']; print('$v', round(p['ms'],3), {k: round(x,3) for k,x in p['kernels_ms'].items() if x>0.05})"
done
