set -x
mkdir -p gpurun_out
rm -f gpurun_out/r02_parity_report.tsv
ONIRIS_PARITY_REPORT=gpurun_out/r02_parity_report.tsv timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02_gputest6.log
tail -12 gpurun_out/r02_gputest6.log
python -c "import __graft_entry__ as g; g.smoke()"
