set -x
mkdir -p gpurun_out
rm -f gpurun_out/vae_parity.tsv
ONIRIS_PARITY_REPORT=gpurun_out/vae_parity.tsv timeout 600 python -m pytest tests/test_vae_gpu.py tests/test_conv_gpu.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r02_gputest_vae.log
tail -12 gpurun_out/r02_gputest_vae.log
grep test_vae gpurun_out/vae_parity.tsv | cut -f1-4 | sed 's/tests.test_vae_gpu.py:://' | head -40
timeout 900 python tools/bench_vae.py > gpurun_out/r02_vae.log 2>&1
tail -3 gpurun_out/r02_vae.log | head -1
