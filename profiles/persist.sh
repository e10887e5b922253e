# persistent-grid co-scheduling experiment: bash profiles/persist.sh
mkdir -p gpurun_out
run() { # z_ctas stream_ctas pad_kb pipelines
  export H2AES_Z_CTAS_PER_SM=$1 H2AES_STREAM_CTAS_PER_SM=$2 H2AES_STREAM_PAD_KB=$3
  timeout 600 python bench.py --no-cpu --no-e2e --no-prover-side --no-perm --no-config4 --no-sharded --steps 8 --pipelines $4 > gpurun_out/persist.json 2> gpurun_out/persist.err || tail -3 gpurun_out/persist.err
  python - <<PY
import json
d = json.load(open("gpurun_out/persist.json"))
ks = {k["kernel"]: k["ms_per_step"] for k in d["kernels"]}
print("z=$1 stream=$2 pad=$3 pipes=$4:", round(d["value"]), "blocks/s", round(d["ms_per_circuit"], 4), "ms/circuit", {k: round(v, 3) for k, v in ks.items() if v > 0.1})
PY
}
run 0 0 0 2
run 4 0 0 2
run 4 4 56 2
run 4 4 56 3
run 3 4 56 2
run 3 5 44 2
run 4 3 75 2
run 4 4 56 4
