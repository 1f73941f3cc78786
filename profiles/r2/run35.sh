mkdir -p gpurun_out/r2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_nohid.csv python profiles/run_profile.py 2 > gpurun_out/r2/launches_nohid.log 2>&1
