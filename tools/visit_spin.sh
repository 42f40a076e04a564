#!/bin/bash
# GPU visit: all-heads site kernel with polling vs suspended barrier waits (APX_SITE_SPIN bit 0 MMA thread, 1 warpgroups, 2 TMA thread).
set -u
mkdir -p gpurun_out
TAG=${1:-r2p}
OUT=gpurun_out/site_spin_${TAG}.txt
: > $OUT
nvidia-smi -L
for c in site_2x1000x256_8+128_f16 site_16x1000x256_8+128_f16 site_2x1000x256_8+32_f16; do
  for sp in 0 1 5 3 7; do
    echo "== $c APX_SITE_SPIN=$sp" >> $OUT
    APX_SITE_SPIN=$sp timeout 120 tools/apx_harness $c 2>&1 | grep -E "OK|FAIL|time per|pass\]|error" >> $OUT
  done
  for sp in 0 1 7; do
    echo "== trace $c APX_SITE_SPIN=$sp" >> $OUT
    APX_SITE_SPIN=$sp APX_SITE_TRACE=1 APX_SITE_CLUSTER=${CL:-0} timeout 120 tools/apx_harness $c 2>&1 | grep -A40 "apx site trace" | head -45 >> $OUT
  done
done
cat $OUT | cut -c1-260
echo "== flash tests"
timeout 600 python -m pytest tests/test_gpu_flash.py -q -m gpu -x 2>&1 | tail -5
