timeout 900 python -m pytest tests/test_elementwise_gpu.py tests/test_vae_gpu.py tests/test_conv_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python tools/bench_vae.py 2>&1 | tail -2 | cut -c1-300
timeout 400 python tools/profile_vae.py 2>&1 | grep -v Warn | sed -n 2,22p
