mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_final.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["gpu_launches"])
print({k:v for k,v in d["roofline"].items() if k in("achieved","frac","achieved_3d_microsteps","frac_3d_microsteps")})
for r in d["secondary"]["config4_attention_microbench"]["rows"]: print(r["seq_len"], round(r["fwd_ms"],3), round(r["bwd_ms"],3), round(r["fwd_tflops_sparse"]), round(r["bwd_tflops_sparse"]), round(r["fwd_frac_of_bf16_burst"],3), round(r["bwd_frac_of_bf16_burst_executed"],3))
PY
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_attention_gpu.py -m gpu -q -x > gpurun_out/sanitizer_attention_v3.log 2>&1; echo "sanitizer rc=$?"; tail -3 gpurun_out/sanitizer_attention_v3.log
