#!/bin/bash
# GPU batch C (N GPUs): sharded step == single-GPU step for both exchanges (+ graph replay), then the N-GPU bench line
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/sp_check.py > gpurun_out/r2_sp_check_n$N.log 2>&1
grep "sp_check\|Error\|error" gpurun_out/r2_sp_check_n$N.log | tail -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 ${BENCH_ARGS} > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 2500 gpurun_out/r2_bench_n$N.json; tail -5 gpurun_out/r2_bench_n$N.err
