# scratch job script for `gpurun -- 'bash tools/gpu_job.sh'`
mkdir -p gpurun_out
for w in mid last; do
NCU_STEP=$w timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_step_$w.csv python tools/ncu_step.py > gpurun_out/ncu_step_$w.log 2>&1; wc -l gpurun_out/r02_launches_step_$w.csv
done
