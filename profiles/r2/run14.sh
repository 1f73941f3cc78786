mkdir -p gpurun_out/r2
export CB200_ACCUM_MODE=4
CB200_EXTRA_NVCC_FLAGS="-DCB_PHASE_TIMING" python -c "from confidence_bootstrapping_b200 import build as b; print(b.build_library(force=True))"
timeout 600 python profiles/phase_timing_ws.py > gpurun_out/r2/phase_ws.log 2>&1
python -c "from confidence_bootstrapping_b200 import build as b; print(b.build_library(force=True))"
cat gpurun_out/r2/phase_ws.log
