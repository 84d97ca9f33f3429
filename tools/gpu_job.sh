set -x
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_report.tsv
ONIRIS_PARITY_REPORT=gpurun_out/r02_parity_report.tsv timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_gputest5.log
tail -25 gpurun_out/r02_gputest5.log
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err
tail -3 gpurun_out/r02_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline']['frac'])
for k in d['roofline_hbm']['kernels']: print(k)
print(d['roofline_hbm']['other_entry_points_ms_per_cycle'], d['roofline_hbm']['serial_cycle_ms'])
PY
