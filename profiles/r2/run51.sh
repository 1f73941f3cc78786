mkdir -p gpurun_out/r2
export PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True
bash profiles/r2/run47.sh
