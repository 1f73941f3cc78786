mkdir -p gpurun_out/r2
export CB200_ACCUM_MODE=4
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_m4.csv python profiles/run_profile.py 2 > gpurun_out/r2/launches_m4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tp_accumulate -s 19 -c 1 -f -o gpurun_out/r2/k3_m4 python profiles/run_profile.py 2 > gpurun_out/r2/k3_m4.log 2>&1
tail -2 gpurun_out/r2/k3_m4.log
