mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_unet_gpu.py tests/test_train_gpu.py tests/test_reference_swap_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 900 python bench.py --steps 40 --warmup 5 --no-secondary --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])
"
