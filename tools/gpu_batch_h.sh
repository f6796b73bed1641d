#!/bin/bash
# GPU batch H (1 GPU): the opt-in three-q-tile attention kernel (variant 0x20): its parity cases, timing against the default
# kernel, and -- only if both are fine -- the whole -m gpu suite with it as the default kernel (PF_TEST_ATTN_TRIPLE=1)
mkdir -p gpurun_out
PF_TEST_ATTN_EXTRA=0x20 timeout 200 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" -p no:cacheprovider > gpurun_out/r2_attn3q_tests.log 2>&1
tail -4 gpurun_out/r2_attn3q_tests.log
if ! grep -q "2 passed" gpurun_out/r2_attn3q_tests.log; then echo "[batch_h] three-q-tile kernel FAILED its parity cases; stopping"; exit 0; fi
PF_CHECK_TIMEOUT=100 timeout 150 python tools/gpu_check.py attn_perf 2>&1 | grep "attn_perf"
PF_SWEEP_VARIANTS="0x20" PF_SWEEP_DELAYS="0 400 800 1200" PF_CHECK_TIMEOUT=100 timeout 150 python tools/gpu_check.py attn_phase_sweep 2>&1 | grep "attn_phase_sweep"
PF_TEST_ATTN_TRIPLE=1 timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -v "^$\|it/s\|Warning\|warnings\|conv.wrap\|frozen" > gpurun_out/r2_tests_triple.log
tail -4 gpurun_out/r2_tests_triple.log
