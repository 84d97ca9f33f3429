mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 900 python bench.py --steps 40 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r02_bench_t.json 2> gpurun_out/r02_bench_t.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_t.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], {k:v for k,v in d["roofline"].items() if k in("achieved","frac","achieved_3d_microsteps","frac_3d_microsteps")})
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_attention_gpu.py -m gpu -q -x > gpurun_out/sanitizer_attention_v2.log 2>&1; echo "sanitizer rc=$?"; tail -5 gpurun_out/sanitizer_attention_v2.log
