#!/bin/bash
# bench step under two settings of one environment switch:  tools/visit_bench_ab.sh TAG "ENV=a" "ENV=b"
set -u
mkdir -p gpurun_out
TAG=$1; shift
i=0
for e in "$@"; do
  i=$((i+1))
  echo "== $e"
  env $e timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err
  python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_${TAG}_$i.json") if l.startswith("{")][-1]
print("value", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), round(d["e2e"]["ms_per_step"],3), "launches/step", d["gpu_launches_per_step"], d["clocks"].get("settle"))
PY
  tail -2 gpurun_out/bench_${TAG}_$i.err
done
