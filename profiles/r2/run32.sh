mkdir -p gpurun_out/r2
ncu --set full --clock-control none --import-source on -k regex:tp_ -s 38 -c 2 -f -o gpurun_out/r2/k3_final python profiles/run_profile.py 2 > gpurun_out/r2/k3_final.log 2>&1
tail -3 gpurun_out/r2/k3_final.log
