# scratch job script for `gpurun -- 'bash tools/gpu_job.sh'`
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 5 2>/dev/null > gpurun_out/r02_bench_n8_final.json
python -c "
import json
for l in open('gpurun_out/r02_bench_n8_final.json'):
    if l.startswith('{'):
        d=json.loads(l); print('N=8', d['value'], d['ms_per_step'], d['e2e']['value'])
"
