#!/bin/bash
# One GPU visit (run under gpurun from the repo root):  tools/gpu_visit.sh TAG [steps...]
#   steps: smoke tests testsall ubench harness bench bench8 ncu_site ncu_all launches profile    (default: smoke tests bench)
set -u
mkdir -p gpurun_out
TAG=${1:-r2}; shift || true
STEPS=${*:-smoke tests bench}
nvidia-smi -L
for s in $STEPS; do
  case $s in
    smoke)   echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 ;;
    tests)   echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 ;;
    testsall) echo "== pytest -m gpu (no -x)"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 ;;
    ubench)  echo "== ubench3"; timeout 120 tools/ubench3 2>&1 | tee gpurun_out/ubench3_${TAG}.txt ;;
    harness) echo "== harness"; timeout 600 tools/apx_harness > gpurun_out/harness_${TAG}.log 2>&1; grep -E "^\[FAIL|time|harness:" gpurun_out/harness_${TAG}.log | tail -60 ;;
    bench)   echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -c 3000 gpurun_out/bench_${TAG}.json; tail -5 gpurun_out/bench_${TAG}.err ;;
    bench8)  echo "== bench batch 8"; timeout 600 python bench.py --batch 8 --steps 50 --no-cpu-baseline > gpurun_out/bench_${TAG}_b8.json 2> gpurun_out/bench_${TAG}_b8.err; tail -c 1500 gpurun_out/bench_${TAG}_b8.json; tail -5 gpurun_out/bench_${TAG}_b8.err ;;
    ncu_site)
      echo "== ncu full (site kernels, harness cases)"
      for c in site_2x1000x256_8+128 site_16x1000x256_8+128 site_2x252x384_8+128 site_2x64x640_8+128; do
        timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ipattn_' -s 3 -c 1 -o gpurun_out/prof_${c}_${TAG} -f \
            tools/apx_harness $c > gpurun_out/ncu_${c}_${TAG}.log 2>&1
      done ;;
    launches)
      echo "== ncu launch list (one eager denoise step)"
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_${TAG}.csv \
          python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_${TAG}.log 2>&1
      wc -l gpurun_out/launches_${TAG}.csv ;;
    ncu_all)
      echo "== ncu full, every libapx kernel once (tools/ncu_kernels.py); the report stays on the box, the raw page comes back"
      timeout 1500 ncu --set full --clock-control none -k regex:'gemm|ipattn|flash|groupnorm|add_bias|kv_pack|layernorm|pool_tokens|patchify|fbank|wave_mean|resample|wgrad' \
          -o /tmp/prof_${TAG}_all -f python tools/ncu_kernels.py > gpurun_out/ncu_${TAG}_all.log 2>&1
      ncu -i /tmp/prof_${TAG}_all.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_all_raw.csv 2>/dev/null
      python tools/ncu_summary.py gpurun_out/prof_${TAG}_all_raw.csv > gpurun_out/ncu_${TAG}_all_kernels.txt; gzip -f gpurun_out/prof_${TAG}_all_raw.csv
      head -70 gpurun_out/ncu_${TAG}_all_kernels.txt | cut -c1-200 ;;
    profile) echo "== step profile"; PYTHONPATH=. timeout 600 python tools/profile_step.py > gpurun_out/step_profile_${TAG}.txt 2>&1; head -40 gpurun_out/step_profile_${TAG}.txt ;;
    *) echo "unknown step $s" ;;
  esac
done
ls -la gpurun_out | tail -15
