#!/bin/bash
# GPU batch B (1 GPU): the whole -m gpu suite + smoke, attention timing / timeline / CTA trace, ncu captures for profiles/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -v "^$\|it/s" > gpurun_out/r2_tests_full.log
tail -6 gpurun_out/r2_tests_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
PF_CHECK_TIMEOUT=200 timeout 300 python tools/gpu_check.py attn_perf 2>&1 | grep "attn_perf"
PF_TL_PHASE="0 800" PF_CHECK_TIMEOUT=100 timeout 150 python tools/gpu_check.py attn_timeline 2>&1 | grep "attn_timeline"
PF_TRACE_VARIANT=0x10 PF_CHECK_TIMEOUT=100 timeout 150 python tools/gpu_check.py attn_cta_trace 2>&1 | grep "attn_cta_trace"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn2_fwd -s 1 -c 1 -o gpurun_out/r2_attn2_final -f python tools/prof_one.py attn 2 0 > gpurun_out/ncu_attn2_final.log 2>&1
tail -2 gpurun_out/ncu_attn2_final.log
timeout 600 ncu --kernel-name-base mangled -k regex:_ZN2pf --metrics gpu__time_duration.sum --clock-control none -c 512 --csv --log-file gpurun_out/r2_step_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu --no-vae --no-video --no-eager > gpurun_out/ncu_step.log 2>&1
tail -c 300 gpurun_out/ncu_step.log; wc -l gpurun_out/r2_step_launches.csv
