#!/bin/bash
# Final-state visit: full GPU suite, smoke, batch-8 step, config 3 / 4 lines.
set -u
mkdir -p gpurun_out
TAG=${1:-r3c}
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench batch 8"; timeout 600 python bench.py --batch 8 --steps 50 --no-cpu-baseline --no-extras > gpurun_out/bench_${TAG}_b8.json 2> gpurun_out/bench_${TAG}_b8.err; tail -2 gpurun_out/bench_${TAG}_b8.err
echo "== config3"; timeout 600 python bench.py --workload config3 > gpurun_out/bench_${TAG}_config3.json 2> gpurun_out/bench_${TAG}_config3.err; tail -c 900 gpurun_out/bench_${TAG}_config3.json; tail -2 gpurun_out/bench_${TAG}_config3.err
echo "== config4"; timeout 600 python bench.py --workload config4 --no-cpu-baseline --no-extras > gpurun_out/bench_${TAG}_config4.json 2> gpurun_out/bench_${TAG}_config4.err; tail -2 gpurun_out/bench_${TAG}_config4.err
python - <<PY
import json
for n in ("b8","config4"):
    try:
        d=[json.loads(l) for l in open("gpurun_out/bench_${TAG}_%s.json"%n) if l.startswith("{")][-1]
        print(n, "value", round(d["value"],2), d["unit"], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "torch", d.get("torch_gpu",{}).get("ms_per_step"), "roofline", d.get("roofline",{}).get("us_per_call"), d.get("roofline",{}).get("frac"))
    except Exception as e: print(n, "failed", e)
PY
