#!/bin/bash
# bench step with the one-head kernel allowed at C = 384 (old policy) vs three kernels there (new default)
set -u
mkdir -p gpurun_out
for m in 320 384; do
  echo "== APX_SITE_H1_MAXC=$m"
  APX_SITE_H1_MAXC=$m timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/bench_r2z_h1max$m.json 2> gpurun_out/bench_r2z_h1max$m.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2z_h1max$m.json"))
print("value", round(d["value"],2), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), round(d["e2e"]["ms_per_step"],3), "launches/step", d["gpu_launches_per_step"], d["clocks"].get("settle"), "torch", d.get("torch_gpu",{}).get("ms_per_step"))
PY
  tail -2 gpurun_out/bench_r2z_h1max$m.err
done
