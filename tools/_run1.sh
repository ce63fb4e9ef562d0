timeout 300 python tools/head_probe.py
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q -s -k "head" 2>&1 | grep -E "head_g|passed|failed" | tail -8
