set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -m gpu -q 2>&1 | tail -3
timeout 600 python tools/bench_attention.py 2>&1 | grep -v "^\[" | head -8
timeout 900 python bench.py --steps 40 --warmup 5 > gpurun_out/r02_bench_16w.json 2> gpurun_out/r02_bench_16w.err; tail -c 3000 gpurun_out/r02_bench_16w.json
