# scratch job script for `gpurun -- 'bash tools/gpu_job.sh'`
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/check_dp_consistency.py 2>&1 | grep -E "world=|digest difference|Error|error|assert" | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 40 --warmup 8 --no-secondary 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'])
"
