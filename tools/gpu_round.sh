#!/bin/bash
# One GPU visit: smoke, GPU parity tests, kernel harness, bench, ncu captures of our kernels.  Run under gpurun from the repo root.
#   tools/gpu_round.sh TAG [full]     "full" adds the (slow) ncu launch list of one bench step
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
nvidia-smi -L
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== harness"; timeout 600 tools/apx_harness > gpurun_out/harness_${TAG}.log 2>&1; grep -E "^\[FAIL|time|harness:" gpurun_out/harness_${TAG}.log | tail -40
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 2500 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
echo "== bench batch 8"; timeout 600 python bench.py --batch 8 --steps 50 --no-cpu-baseline > gpurun_out/bench_${TAG}_b8.json 2> gpurun_out/bench_${TAG}_b8.err; tail -c 1200 gpurun_out/bench_${TAG}_b8.json; tail -5 gpurun_out/bench_${TAG}_b8.err
echo "== ncu full (our kernels, harness cases)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ipattn_site' -s 3 -c 1 -o gpurun_out/prof_site_b1_${TAG} -f \
    tools/apx_harness site_2x1000x256_8+128 > gpurun_out/ncu_site_b1_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ipattn_site' -s 3 -c 1 -o gpurun_out/prof_site_b8_${TAG} -f \
    tools/apx_harness site_16x1000x256_8+128 > gpurun_out/ncu_site_b8_${TAG}.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc' -s 2 -c 1 -o gpurun_out/prof_gemm_splitk_${TAG} -f \
    tools/apx_harness gemm_126x640x2560 > gpurun_out/ncu_gemm_splitk_${TAG}.log 2>&1
if [ "${2:-}" = "full" ]; then
  echo "== ncu launch list (one eager denoise step)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_${TAG}.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
  wc -l gpurun_out/launches_${TAG}.csv
fi
ls -la gpurun_out | tail -12
