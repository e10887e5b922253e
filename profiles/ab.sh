# A/B of library variants: bash profiles/ab.sh default v1 v2 ...  (variants under lib/variants)
mkdir -p gpurun_out
for v in "$@"; do
  if [ $v = default ]; then unset H2AES_LIB; else export H2AES_LIB=$PWD/halo2_aes_b200/lib/variants/libh2aes_$v.so; fi
  timeout 600 python bench.py --no-cpu --no-e2e --no-prover-side --no-perm ${AB_ARGS} --steps 100 > gpurun_out/ab_$v.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
d = json.load(open("gpurun_out/ab_$v.json"))
ks = {k["kernel"]: k["ms_per_step"] for k in d["kernels"]}
print("$v", round(d["value"]), round(d["ms_per_step"], 3), "iso-sum", round(sum(d["phase_ms_isolated"].values()), 3),
      {k: round(v, 3) for k, v in ks.items() if v > 0.04})
PY
done
unset H2AES_LIB
