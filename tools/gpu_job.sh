mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_vae_gpu.py -m gpu -q -x 2>&1 | tail -2
cp autoregressive_diffusion_b200/liboniris_b200.so build/variants/lib_main.so
for v in old new old new; do
  cp build/variants/lib_$v.so autoregressive_diffusion_b200/liboniris_b200.so
  echo "== $v"
  timeout 600 python bench_vae_run.py 2>/dev/null | tail -1
  timeout 900 python bench.py --steps 40 --warmup 5 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'])
"
done
cp build/variants/lib_main.so autoregressive_diffusion_b200/liboniris_b200.so
