#!/bin/bash
# GPU visit for the fused kernel: standalone harness (forked cases), GPU parity tests, optional ncu capture.
set -u
mkdir -p gpurun_out
TAG=${1:-r1b}
FILT=${2:-site}
nvidia-smi -L
echo "== harness ($FILT)"; timeout 600 tools/apx_harness $FILT > gpurun_out/harness_${TAG}.log 2>&1; grep -E "^\[|time per|FAIL|product path|harness:" gpurun_out/harness_${TAG}.log | tail -120
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25
if [ "${3:-}" = "ncu" ]; then
  echo "== ncu full: fused kernel at Bc=16 (harness case site_16x1000x256)"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ipattn_fused -s 2 -c 2 -o gpurun_out/prof_fused_${TAG} -f \
      tools/apx_harness site_16x1000x256 > gpurun_out/ncu_fused_${TAG}.log 2>&1
  tail -5 gpurun_out/ncu_fused_${TAG}.log
fi
