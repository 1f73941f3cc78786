mkdir -p gpurun_out/r2
(time python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_fullsize.py::test_config1_1a0q_sampling_and_confidence_vs_oracle --deselect tests/test_gpu_fullsize.py::test_config5_slice_all_atom_score_model_vs_oracle --durations=8) > gpurun_out/r2/t_all3.log 2>&1
tail -15 gpurun_out/r2/t_all3.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config3 > gpurun_out/r2/bench_run3.json 2> gpurun_out/r2/bench_run3.err
cut -c1-300 gpurun_out/r2/bench_run3.json; tail -3 gpurun_out/r2/bench_run3.err
