mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_unet_gpu.py tests/test_vae_gpu.py -m gpu -q -x 2>&1 | tail -3
cp autoregressive_diffusion_b200/liboniris_b200.so build/variants/lib_main.so
for v in old e8 old e8; do
  cp build/variants/lib_$v.so autoregressive_diffusion_b200/liboniris_b200.so
  echo "== $v"
  timeout 600 python tools/conv_breakdown.py > gpurun_out/breakdown_$v.txt 2>&1; head -1 gpurun_out/breakdown_$v.txt
done
cp build/variants/lib_main.so autoregressive_diffusion_b200/liboniris_b200.so
