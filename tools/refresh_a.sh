# round-2 refresh, part A: full GPU test-suite, headline bench, datasets, 1 M-pose grid, DGEMM witness
timeout 700 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_gputest_final.log
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python tools/dgemm_peak.py > gpurun_out/r2_dgemm_witness.json 2> gpurun_out/r2_dgemm_witness.err
python tools/dataset_bench.py > gpurun_out/r2_datasets.json 2> gpurun_out/r2_datasets.err
SPG_HOST_PROF=1 python tools/grid_bench.py --rows 1000 --cols 1000 > gpurun_out/r2_grid_1000.json 2> gpurun_out/r2_grid_1000.err
cat gpurun_out/r2_gputest_final.log; cut -c1-200 gpurun_out/r2_bench.json; cat gpurun_out/r2_dgemm_witness.json; cut -c1-300 gpurun_out/r2_grid_1000.json
