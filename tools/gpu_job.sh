set -x
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q 2>&1 | tail -3
timeout 600 python tools/bench_attention.py 2>&1 | grep -v "^\[" | head -8
