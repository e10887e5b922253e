mkdir -p gpurun_out
for cfg in "0 default 2" "56 default 2" "56 default 3" "56 occ4 2" "56 occ4 3" "75 occ4 3" "40 occ4 3"; do
  set -- $cfg
  export H2AES_STREAM_PAD_KB=$1
  if [ $2 = default ]; then unset H2AES_LIB; else export H2AES_LIB=$PWD/halo2_aes_b200/lib/variants/libh2aes_$2.so; fi
  timeout 600 python bench.py --no-cpu --no-e2e --no-prover-side --no-perm --steps 10 --pipelines $3 > gpurun_out/pad.json 2> gpurun_out/pad.err || tail -3 gpurun_out/pad.err
  python - <<PY
import json
d = json.load(open("gpurun_out/pad.json"))
ks = {k["kernel"]: k["ms_per_step"] for k in d["kernels"]}
print("pad $1 lib $2 pipes $3:", round(d["value"]), "blocks/s", round(d["ms_per_circuit"], 4), "ms/circuit", {k: round(v, 3) for k, v in ks.items() if v > 0.09})
PY
done
