set -x
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_report.tsv
ONIRIS_PARITY_REPORT=gpurun_out/r02_parity_report.tsv timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r02_gputest2.log
tail -40 gpurun_out/r02_gputest2.log
