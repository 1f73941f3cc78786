mkdir -p gpurun_out/r2
(timeout 900 python -m pytest tests/test_gpu_model.py -q -m gpu -x) > gpurun_out/r2/t_54.log 2>&1
tail -3 gpurun_out/r2/t_54.log
timeout 600 python profiles/e2e_profile.py 12 > gpurun_out/r2/e2e_profile2.log 2>&1
head -3 gpurun_out/r2/e2e_profile2.log
bash profiles/r2/run47.sh
