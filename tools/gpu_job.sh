cp autoregressive_diffusion_b200/liboniris_b200.so build/variants/lib_main.so
for v in main pre main pre; do
  cp build/variants/lib_$v.so autoregressive_diffusion_b200/liboniris_b200.so
  echo "== $v"
  timeout 600 python -c "
import sys; sys.path.insert(0,'tools')
import bench_attention as b
for n,hw in ((128,256),(256,256)):
    r=b.run(n,hw,reps=9); print('L=%d fwd %.3f ms %.0f TF   bwd %.3f ms %.0f TF' % (r['seq_len'], r['fwd_ms'], r['fwd_tflops_sparse'], r['bwd_ms'], r['bwd_tflops_sparse']))
"
done
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q -x 2>&1 | tail -2
cp build/variants/lib_main.so autoregressive_diffusion_b200/liboniris_b200.so
