mkdir -p gpurun_out
for c in 4 8 16; do
NCCL_MAX_CTAS=$c timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 8 --steps 20 --warmup 5 --no-secondary 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('max_ctas=$c', d['value'], d['ms_per_step'], d['e2e']['value'])
"
done
