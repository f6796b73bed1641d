python -m pytest tests -m gpu -x -q 2>&1 | tail -4
PF_CONV_KWREUSE=0 python -m pytest tests/test_vae_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/gpu_check.py vae_perf 2>&1 | grep "^\["
