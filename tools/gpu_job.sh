mkdir -p gpurun_out; rm -f gpurun_out/r02_parity_report.tsv
ONIRIS_PARITY_REPORT=$PWD/gpurun_out/r02_parity_report.tsv timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
wc -l gpurun_out/r02_parity_report.tsv
