# scratch job script for `gpurun -- 'bash tools/gpu_job.sh'`: the full GPU validation
mkdir -p gpurun_out
for i in 1 2 3; do timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -1; done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
