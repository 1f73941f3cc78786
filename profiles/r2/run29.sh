mkdir -p gpurun_out/r2
export CB200_ACCUM_MODE=4
timeout 600 python profiles/phase_timing_ws.py > gpurun_out/r2/phase_ws7.log 2>&1
tail -70 gpurun_out/r2/phase_ws7.log
