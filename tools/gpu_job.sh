set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_swap_gpu.py tests/test_vae_gpu.py -m gpu -q -s 2>&1 | tail -30 > gpurun_out/r02_gputest_swap.log
tail -30 gpurun_out/r02_gputest_swap.log
