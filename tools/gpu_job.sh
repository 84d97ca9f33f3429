set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 40 --warmup 5 > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err
tail -3 gpurun_out/r02_bench4.err
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"tapconv_kernel<64, 128, 0, 1>" -s 8 -c 4 -f -o gpurun_out/r02_prof_tapconv_pair python tools/ncu_step.py > /dev/null 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"wnorm_fwd_multi|adamw" -c 1 -f -o gpurun_out/r02_prof_wnorm_multi env NCU_STEP=first python tools/ncu_step.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
