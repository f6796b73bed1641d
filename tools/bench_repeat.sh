for i in 1 2 3; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-vae --no-video --no-eager 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('[bench_rep] run $i: graph ms %.2f | host-launched ms %.2f | e2e ms %.2f | attn %.3f | clocks %s' % (d['ms_per_step'], d['config']['ms_per_step_host_launched'], d['e2e']['ms_per_step'], d['roofline']['avg_launch_ms'], d['clocks']))
"
done
timeout 300 python bench.py --steps 10 --warmup 12 --no-cpu --no-vae --no-video --no-eager 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('[bench_rep] warmup 12: graph ms %.2f | host-launched ms %.2f | e2e ms %.2f | attn %.3f | clocks %s' % (d['ms_per_step'], d['config']['ms_per_step_host_launched'], d['e2e']['ms_per_step'], d['roofline']['avg_launch_ms'], d['clocks']))
"
