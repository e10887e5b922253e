# circuits in flight per GPU: bash profiles/pipes.sh 1 2 3 4
mkdir -p gpurun_out
for p in "$@"; do
  timeout 200 python bench.py --no-cpu --no-e2e --no-prover-side --no-perm --no-config4 --no-sharded --no-check --steps 20 --pipelines $p > gpurun_out/pipes_$p.json 2> gpurun_out/pipes.err || tail -3 gpurun_out/pipes.err
  python - <<PY
import json
d = json.load(open("gpurun_out/pipes_$p.json"))
print("pipelines $p", round(d["value"]), "blocks/s", round(d["ms_per_circuit"], 4), "ms/circuit", {k: round(v, 3) for k, v in d["phase_ms"].items()})
PY
done
