#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== flash default (try_wait)"; PYTHONPATH=.:tools timeout 300 python tools/time_flash.py 2>&1 | tail -6 | tee gpurun_out/flash_ab_r2w.txt
cp ap_adapter_b200/libapx.so /tmp/libapx_main.so; cp variants/ftw/libapx.so ap_adapter_b200/libapx.so
echo "== flash test_wait"; PYTHONPATH=.:tools timeout 300 python tools/time_flash.py 2>&1 | tail -6 | tee -a gpurun_out/flash_ab_r2w.txt
timeout 300 python -m pytest tests/test_gpu_flash.py -q -m gpu -x 2>&1 | tail -3
cp /tmp/libapx_main.so ap_adapter_b200/libapx.so
