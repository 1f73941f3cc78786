# launch list (every kernel of 2 forwards of the bench batch) + full-set capture of the K3 pair of the SECOND conv layer
# (74->74, 4 slots, all 17 600 aggregation nodes; conv layer 0 shares its rec->rec slot across the samples since round 2)
mkdir -p gpurun_out/r2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches.csv python profiles/run_profile.py 2 > gpurun_out/r2/launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tp_ -s 38 -c 2 -f -o gpurun_out/r2/k3 python profiles/run_profile.py 2 > gpurun_out/r2/k3.log 2>&1
tail -3 gpurun_out/r2/k3.log
