#!/bin/bash
# One GPU visit: smoke, GPU parity tests, bench, ncu launch list + full capture of our kernels.  Run under gpurun from the repo root.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
nvidia-smi -L
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err
echo "== bench batch 8"; timeout 600 python bench.py --batch 8 --steps 50 --no-cpu-baseline > gpurun_out/bench_${TAG}_b8.json 2> gpurun_out/bench_${TAG}_b8.err; tail -c 3000 gpurun_out/bench_${TAG}_b8.json; tail -5 gpurun_out/bench_${TAG}_b8.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
wc -l gpurun_out/launches_${TAG}.csv
echo "== ncu full (our kernels)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ipattn|gemm_tc' -s 200 -c 6 -o gpurun_out/prof_${TAG} -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -12
