#!/bin/bash
# GPU batch E (1 GPU): attention variants: correctness, timing, per-iteration timeline and CTA phase trace of the pipelined kernel
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/r2_attn5_tests.log 2>&1
tail -5 gpurun_out/r2_attn5_tests.log
PF_CHECK_TIMEOUT=200 timeout 300 python tools/gpu_check.py attn_perf 2>&1 | grep "attn_perf"
PF_TL_VARIANTS="${TL_VARIANTS:-0x08}" PF_CHECK_TIMEOUT=100 timeout 150 python tools/gpu_check.py attn4_timeline 2>&1 | grep "attn4_timeline" | grep -v " j=1[6-9] \| j=2[01] "
PF_TL_VARIANTS="${TL_VARIANTS:-0x08}" PF_CHECK_TIMEOUT=100 timeout 150 python tools/gpu_check.py attn4_timeline 2>&1 | grep " j=1[67] "
for v in ${TRACE_VARIANTS:-0x08}; do
  PF_TRACE_VARIANT=$v PF_CHECK_TIMEOUT=100 timeout 150 python tools/gpu_check.py attn_cta_trace 2>&1 | grep "attn_cta_trace" | head -8 | sed "s/^/[variant $v] /"
done
