#!/bin/bash
# GPU batch G (1 GPU): LEAN bits x tile phase (isolated launch), then the same choices inside the DiT step (short bench lines)
mkdir -p gpurun_out
PF_SWEEP_VARIANTS="0x30 0x0d 0x0e 0x0c" PF_SWEEP_DELAYS="0 800" PF_CHECK_TIMEOUT=300 timeout 400 python tools/gpu_check.py attn_phase_sweep 2>&1 | grep "attn_phase_sweep\|Error\|error"
for cfg in "0 0" "0 800" "0x0d 0" "0x0d 800"; do
  set -- $cfg
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-vae --no-video --no-eager --attn-variant $1 --attn-phase $2 2> gpurun_out/r2_bench_ab.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('[bench_ab] variant $1 phase $2: ms_per_step %.2f, attention avg_launch_ms %.3f, frac %.3f, sm_mhz %s' % (d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['clocks'].get('sm_mhz')))
"
done
tail -3 gpurun_out/r2_bench_ab.err
