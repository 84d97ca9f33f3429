mkdir -p gpurun_out
cp autoregressive_diffusion_b200/liboniris_b200.so /tmp/lib_main.so
for v in v0 v1 v2 v3; do
  cp build/variants/lib_$v.so autoregressive_diffusion_b200/liboniris_b200.so
  echo "== $v"
  timeout 600 python tools/bench_attention.py 2>&1 | grep -v "^\[" | grep -E "'seq_len': (8192|65536|131072)" | sed -E "s/.*'tokens_per_frame': ([0-9]+), 'seq_len': ([0-9]+).*'fwd_ms': ([0-9.]+).*'fwd_tflops_sparse': ([0-9.]+).*/hw=\1 L=\2 fwd_ms=\3 TF=\4/"
done
cp /tmp/lib_main.so autoregressive_diffusion_b200/liboniris_b200.so
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q 2>&1 | tail -3
