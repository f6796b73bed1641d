python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/gpu_check.py gemm_qkv_perf attn_perf 2>&1 | grep "^\[" 
python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_r1_h.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1_h.json').read()); print(d['ms_per_step'], d['value'], d['clocks'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['ms_per_step'], d['config']['ms_per_step_host_launched'], d['config'].get('breakdown_ms_one_step'))"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd -s 1 -c 1 -f -o gpurun_out/r01_attn_v2 python tools/prof_one.py attn 2 > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -s 1 -c 1 -f -o gpurun_out/r01_gemm2 python tools/prof_one.py gemm 2 > gpurun_out/ncu_gemm.log 2>&1; tail -2 gpurun_out/ncu_gemm.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3d -s 1 -c 1 -f -o gpurun_out/r01_conv python tools/prof_one.py conv 2 > gpurun_out/ncu_conv.log 2>&1; tail -2 gpurun_out/ncu_conv.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm|attn_fwd|ln_modulate|small_linear|patchify|timestep_emb" -c 3200 --csv --log-file gpurun_out/r01_step_launches_v2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-graph > gpurun_out/ncu_step.log 2>&1; tail -c 300 gpurun_out/ncu_step.log; wc -l gpurun_out/r01_step_launches_v2.csv
