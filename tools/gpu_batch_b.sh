#!/bin/bash
# GPU batch B (1 GPU): the whole -m gpu suite, then the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider 2>&1 | grep -v "^$\|it/s" > gpurun_out/r2_tests_full.log
tail -25 gpurun_out/r2_tests_full.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -c 3000 gpurun_out/r2_bench_n1.json; tail -5 gpurun_out/r2_bench_n1.err
