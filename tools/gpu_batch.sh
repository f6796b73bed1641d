python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/gpu_check.py vae_perf 2>&1 | grep "^\["
timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv3d2w -s 1 -c 1 -f -o gpurun_out/r01_conv_kw python tools/prof_one.py conv 2 > gpurun_out/ncu_conv_kw.log 2>&1; tail -2 gpurun_out/ncu_conv_kw.log
