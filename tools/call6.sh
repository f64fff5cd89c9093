mkdir -p gpurun_out
python tools/prof_tiled.py > gpurun_out/s2_prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --launch-skip 2 -c 2 -k regex:gine_aggr_.*tiled -f -o gpurun_out/r02_tiled_final python tools/prof_tiled.py > gpurun_out/s2_ncu_tiled.log 2>&1
python bench.py --steps 2 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/s2_bench_short.json 2> gpurun_out/s2_bench_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-roofline --no-cpu-baseline > gpurun_out/s2_ncu_bench.log 2>&1
ls -la gpurun_out/r02_tiled_final.ncu-rep gpurun_out/r02_launches_bench_final.csv
