#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== flash, exp2 token (default)"; PYTHONPATH=.:tools timeout 300 python tools/time_flash.py 2>&1 | tail -5 | tee gpurun_out/flash_tok_ab_r3f.txt
cp ap_adapter_b200/libapx.so /tmp/libapx_main.so; cp variants/fnotok/libapx.so ap_adapter_b200/libapx.so
echo "== flash, no token"; PYTHONPATH=.:tools timeout 300 python tools/time_flash.py 2>&1 | tail -5 | tee -a gpurun_out/flash_tok_ab_r3f.txt
timeout 300 python -m pytest tests/test_gpu_flash.py -q -m gpu -x 2>&1 | tail -2
cp /tmp/libapx_main.so ap_adapter_b200/libapx.so
