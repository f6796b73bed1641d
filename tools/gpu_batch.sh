python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/gpu_check.py epilogue gemm_qkv_perf 2>&1 | grep "^\["
python bench.py --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_r1_i.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1_i.json').read()); print(d['ms_per_step'], d['value'], d['clocks'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['ms_per_step'], d['config']['ms_per_step_host_launched'], d['config'].get('breakdown_ms_one_step'))"
timeout 600 python tools/e2e_generate.py --height 384 --width 640 --temp 16 2>&1 | tail -4
