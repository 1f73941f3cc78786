mkdir -p gpurun_out/r2
export CB200_ACCUM_MODE=4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tp_accumulate -s 19 -c 1 -f -o gpurun_out/r2/k3_m4f python profiles/run_profile.py 2 > gpurun_out/r2/k3_m4f.log 2>&1
tail -2 gpurun_out/r2/k3_m4f.log
