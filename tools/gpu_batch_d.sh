#!/bin/bash
# GPU batch D (1 GPU): exponential-stream probe, correctness of every attention variant, attention timing, CTA phase traces
mkdir -p gpurun_out
timeout 120 tools/probes/exp_sched_probe > gpurun_out/r2_exp_sched_probe.txt 2>&1
cat gpurun_out/r2_exp_sched_probe.txt
timeout 400 python -m pytest tests/test_kernels_gpu.py -x -q -k "attention" > gpurun_out/r2_attn4_tests.log 2>&1
tail -5 gpurun_out/r2_attn4_tests.log
PF_CHECK_TIMEOUT=200 timeout 300 python tools/gpu_check.py attn_perf 2>&1 | grep "attn_perf"
for v in ${TRACE_VARIANTS:-0x80 0x83}; do
  PF_TRACE_VARIANT=$v PF_CHECK_TIMEOUT=100 timeout 150 python tools/gpu_check.py attn_cta_trace 2>&1 | grep "attn_cta_trace" | head -8 | sed "s/^/[variant $v] /"
done
