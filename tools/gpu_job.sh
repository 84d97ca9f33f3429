mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 7 python -m pytest tests/test_conv_gpu.py tests/test_elementwise_gpu.py tests/test_vae_gpu.py tests/test_optimizer_gpu.py -m gpu -q -x > gpurun_out/sanitizer_conv_final.log 2>&1; echo "sanitizer rc=$?"; tail -4 gpurun_out/sanitizer_conv_final.log
