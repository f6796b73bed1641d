python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py 2>&1 | tail -1 > gpurun_out/bench_r1_final.json
python -c "
import json; d=json.loads(open('gpurun_out/bench_r1_final.json').read()); print(d['ms_per_step'], d['value'], d['clocks'], d['roofline']['avg_launch_ms'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['ms_per_step'], d['cpu_baseline'])"
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_r1_ref.json; head -c 600 gpurun_out/bench_r1_ref.json
