#!/bin/bash
# compute-sanitizer over the torch-free harness (tools/apx_harness, built by __graft_entry__.build()):
#   tools/sanitize.sh [memcheck|racecheck|synccheck] [case filter ...]      e.g.  tools/sanitize.sh memcheck site_1x200x256 gemm_100x64
# One process per filter so that a slow case cannot eat the others' time; every case compares the tcgen05 kernels with the CUDA-core
# cross-check (and a CPU reference for the *_cpu cases) while the tool watches.  Output: gpurun_out/sanitize_<tool>_<filter>.txt;
# exit code = number of filters with a violation or a failed comparison.  Under the tool the harness's timing loops run 10-100x
# slower: pick the small *_cpu cases (profiles/r3r_compute_sanitizer_memcheck.txt took 20 s per site case).
set -u
TOOL=${1:-memcheck}; shift || true
FILTERS=${*:-"gemm_100x64x1024 gemm_130x128x512 site_1x200x256_8+32 site_1x130x384 site_1x64x640 site_3x129x256 site_1x128x256_8+512"}
mkdir -p gpurun_out
bad=0
for f in $FILTERS; do
  out="gpurun_out/sanitize_${TOOL}_$(echo "$f" | tr -c 'A-Za-z0-9_\n' '_').txt"
  timeout "${SANITIZE_TIMEOUT:-300}" stdbuf -oL compute-sanitizer --tool "$TOOL" --print-limit 20 --error-exitcode 9 tools/apx_harness "$f" > "$out" 2>&1
  rc=$?
  summary=$(grep -E "ERROR SUMMARY" "$out" | tail -1)
  fails=$(grep -c "^\[FAIL" "$out")
  echo "$TOOL $f: rc=$rc fails=$fails ${summary:-"(no summary: timed out or crashed)"}"
  if [ "$rc" -ne 0 ] || [ "$fails" -ne 0 ]; then bad=$((bad + 1)); fi
done
exit $bad
