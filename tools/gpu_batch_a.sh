#!/bin/bash
# GPU batch A (1 GPU): attention kernels (correctness over all variants, then the perf sweep) and the drop-in test
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attention" -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/r2_attn_tests.log
cat gpurun_out/r2_attn_tests.log
PF_CHECK_TIMEOUT=150 python tools/gpu_check.py attn_perf 2>&1 | tail -12
timeout 240 python -m pytest tests/test_dropin_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -v "it/s\|^$" | tail -30 > gpurun_out/r2_dropin.log
cat gpurun_out/r2_dropin.log
