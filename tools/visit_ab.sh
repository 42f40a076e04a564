#!/bin/bash
# GPU visit: A/B runs of harness cases under environment switches / library variants (the A/B logs under profiles/r2p..r3a were made
# with it).  usage: tools/visit_ab.sh TAG < lines of "ENV=.. LD_LIBRARY_PATH=variants/NAME | case-filter | trace?"
# (variants/NAME/libapx.so: tools/build_variant.sh NAME file.cu "-DMACRO=1")
set -u
mkdir -p gpurun_out
TAG=${1:-ab}
OUT=gpurun_out/ab_${TAG}.txt
: > $OUT
nvidia-smi -L >> $OUT
while IFS='|' read -r envs filt trace; do
  [ -z "$filt" ] && continue
  filt=$(echo $filt); trace=$(echo $trace)
  echo "== [$envs] $filt" >> $OUT
  if [ "$trace" = "trace" ]; then
    env $envs APX_SITE_TRACE=1 timeout 120 tools/apx_harness $filt 2>&1 | grep -A22 "apx site trace" | head -48 >> $OUT
  else
    env $envs timeout 180 tools/apx_harness $filt 2>&1 | grep -E "OK|FAIL|time per|time:|pass\]|error|trap|killed" >> $OUT
  fi
done
cut -c1-230 $OUT
