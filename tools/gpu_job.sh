set -x
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"tapconv_kernel<\(int\)64, \(int\)128, \(bool\)0, \(bool\)1>" -s 8 -c 4 -f -o gpurun_out/r02_prof_tapconv_pair python tools/ncu_step.py > gpurun_out/ncu_tapconv.log 2>&1
tail -3 gpurun_out/ncu_tapconv.log
ls -la gpurun_out/*.ncu-rep
