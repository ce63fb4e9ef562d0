timeout 300 python tools/head_probe.py
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q -k "head" 2>&1 | tail -2
