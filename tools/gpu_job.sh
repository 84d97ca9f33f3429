set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q -x 2>&1 | tail -15
timeout 600 python tools/bench_attention.py 2>&1 | grep -v "^\[" | head -8
