# quick GPU check used during optimisation: parity subset + bench with live per-kernel times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} > gpurun_out/pytest.log 2>&1; echo pytest=$?; tail -2 gpurun_out/pytest.log
timeout 600 python bench.py --no-cpu --no-e2e --no-prover-side ${BENCH_ARGS} > gpurun_out/quick.json 2> gpurun_out/quick.err; echo bench=$?; tail -3 gpurun_out/quick.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/quick.json"))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "clocks", d["clocks"])
print("phases(iso)", {k: round(v, 3) for k, v in d["phase_ms_isolated"].items()}, "sum", round(sum(d["phase_ms_isolated"].values()), 3))
for k in d["kernels"]:
    if k["ms_per_step"] > 0.03:
        print(f"  {k['kernel']:18s} {k['ms_per_step']:.4f} ms", f"frac {k['frac']:.3f}" if "frac" in k else "")
PY
