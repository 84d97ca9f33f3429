timeout 900 python -m pytest tests/test_attention_gpu.py -m gpu -q -x 2>&1 | tail -12
timeout 600 python tools/bench_attention.py 2>&1 | grep -v "^\[" | sed -E "s/.*'tokens_per_frame': ([0-9]+), 'seq_len': ([0-9]+).*'fwd_ms': ([0-9.]+), 'bwd_ms': ([0-9.]+), 'fwd_tflops_sparse': ([0-9.]+), 'bwd_tflops_sparse': ([0-9.]+).*/hw=\1 L=\2 fwd_ms=\3 bwd_ms=\4 fwdTF=\5 bwdTF=\6/" | head -6
