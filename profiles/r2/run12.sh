mkdir -p gpurun_out/r2
export CB200_ACCUM_MODE=4
CB200_EXTRA_NVCC_FLAGS="-DCB_WS_NO_STORE" python -c "from confidence_bootstrapping_b200 import build as b; print(b.build_library(force=True))"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/launches_m4_nostore.csv python profiles/run_profile.py 2 > gpurun_out/r2/launches_m4_nostore.log 2>&1
python -c "from confidence_bootstrapping_b200 import build as b; print(b.build_library(force=True))"
