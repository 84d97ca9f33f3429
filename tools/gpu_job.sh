cp autoregressive_diffusion_b200/liboniris_b200.so build/variants/lib_main.so
for v in main q25 main q25 main; do
  cp build/variants/lib_$v.so autoregressive_diffusion_b200/liboniris_b200.so
  echo "== $v"
  timeout 600 python tools/bench_attention.py 2>&1 | grep -v "^\[" | grep -E "'seq_len': (65536|131072)" | sed -E "s/.*'tokens_per_frame': ([0-9]+), 'seq_len': ([0-9]+).*'fwd_ms': ([0-9.]+), 'bwd_ms': ([0-9.]+), 'fwd_tflops_sparse': ([0-9.]+), 'bwd_tflops_sparse': ([0-9.]+).*/L=\2 fwd_ms=\3 bwd_ms=\4 fwdTF=\5 bwdTF=\6/"
  nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv,noheader
done
cp build/variants/lib_main.so autoregressive_diffusion_b200/liboniris_b200.so
