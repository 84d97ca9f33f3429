mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_ -c 12 -f -o gpurun_out/r02_prof_attn_v2 python tools/ncu_attention.py > gpurun_out/ncu_attn_v2.log 2>&1
tail -2 gpurun_out/ncu_attn_v2.log
