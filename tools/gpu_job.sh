timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q -x -k full_size 2>&1 | tail -12
