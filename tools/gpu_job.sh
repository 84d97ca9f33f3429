# scratch job script for `gpurun -- 'bash tools/gpu_job.sh'`: the full GPU validation
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_final.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["gpu_launches"])
print({k:v for k,v in d["roofline"].items() if k in("achieved","frac","achieved_3d_microsteps","frac_3d_microsteps","share_of_step")})
print([ (k["kernel"], k["frac"]) for k in d["roofline_hbm"]["kernels"][:4]])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2>/dev/null; head -c 300 gpurun_out/r02_bench_ref.json
