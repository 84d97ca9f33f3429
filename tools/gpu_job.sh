set -x
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_report.tsv
ONIRIS_PARITY_REPORT=gpurun_out/r02_parity_report.tsv timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -80 > gpurun_out/r02_gputest3.log
tail -50 gpurun_out/r02_gputest3.log
NO_CPU=1 N_GEN=8 timeout 900 python tools/bench_sampling.py > gpurun_out/r02_sampling.log 2>&1
tail -12 gpurun_out/r02_sampling.log
