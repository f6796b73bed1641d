python -m pytest tests/test_dit_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/e2e_generate.py --height 384 --width 640 --temp 16 2>&1 | tail -1
timeout 900 python tools/e2e_generate.py --height 768 --width 1280 --temp 16 2>&1 | tail -1
timeout 1200 python tools/e2e_generate.py --height 768 --width 1280 --temp 31 2>&1 | tail -1
