#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 60 tools/ubench4 | tee gpurun_out/ubench4_r3d.txt
echo "== gemm16 compact no-residual epilogue (new)"; PYTHONPATH=.:tools timeout 300 python tools/time_gemm_step.py 2>&1 | tail -19 | tee gpurun_out/g16_ab_r3e.txt
cp ap_adapter_b200/libapx.so /tmp/libapx_main.so; cp variants/g16old/libapx.so ap_adapter_b200/libapx.so
echo "== gemm16 unrolled epilogue (old)"; PYTHONPATH=.:tools timeout 300 python tools/time_gemm_step.py 2>&1 | tail -19 | tee -a gpurun_out/g16_ab_r3e.txt
cp /tmp/libapx_main.so ap_adapter_b200/libapx.so
timeout 300 python -m pytest tests/test_gpu_ff.py tests/test_gpu_norm.py -q -m gpu -x 2>&1 | tail -3
