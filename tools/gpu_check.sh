#!/bin/bash
# one GPU-box round trip: GPU parity tests, then the default bench; summaries under gpurun_out/
tag=${1:-check}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$tag.json"))
print(round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["pipeline"])
print({k: round(v["ms"], 4) for k, v in d["stages"].items()})
print(d["roofline"])
print(d["extra"])
print(d["cpu_baseline"], d["cpu_baseline_reference"])
PY
cat gpurun_out/pytest_$tag.log; tail -5 gpurun_out/bench_$tag.err
