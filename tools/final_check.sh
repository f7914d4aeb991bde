#!/bin/bash
# End-of-round run on one GPU box: GPU parity tests, smoke, the bench line of every pipeline and the reference arm.
# (The ncu captures are separate calls: tools/ncu_capture.sh for `--set full`, and the launch list
#  `ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv ... python bench.py --steps 2 --warmup 3 --no-extra`.)
tag=${1:-final}
bash tools/gpu_check.sh $tag
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err
for p in fused staged; do
    python bench.py --pipeline $p --no-extra > gpurun_out/bench_${tag}_$p.json 2> gpurun_out/bench_${tag}_$p.err
done
python - <<PY
import json
for n in ("fused", "staged", "ref"):
    d = json.load(open("gpurun_out/bench_${tag}_%s.json" % n))
    print(n, round(d["value"]), round(d["e2e"]["value"]))
PY
