mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 12 -f -o gpurun_out/r02_prof_attn_v3 python tools/ncu_attention.py > gpurun_out/ncu_attn_v3.log 2>&1
tail -1 gpurun_out/ncu_attn_v3.log
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
