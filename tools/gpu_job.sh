# scratch job script for `gpurun -- 'bash tools/gpu_job.sh'`
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -1
timeout 120 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r02_bench_quick.json 2>/dev/null; python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_quick.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac"])
PY
