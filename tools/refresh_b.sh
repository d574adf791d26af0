# round-2 refresh, part B: ncu launch list of the bench command and counters of the small-blanket kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --blankets 20000 --no-cpu-baseline > gpurun_out/r2_launches.log 2>&1
bash tools/ncu_ifetch.sh
cp gpurun_out/ncu_ifetch.csv gpurun_out/r2_ncu_small_aligned.csv
SPG_FAST_WPC=1 bash tools/ncu_ifetch.sh
cp gpurun_out/ncu_ifetch.csv gpurun_out/r2_ncu_small_onewarp.csv
python tools/stage_profile.py 4 5 8 16 > gpurun_out/r2_stage_fast_final.txt 2>&1
wc -l gpurun_out/r2_launches.csv gpurun_out/r2_ncu_small_aligned.csv gpurun_out/r2_ncu_small_onewarp.csv
