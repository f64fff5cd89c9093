# End-of-round check on one GPU: GPU test suite, smoke(), the bench line (both arms).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s2_t_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_t_final.log
tail -3 gpurun_out/s2_t_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s2_smoke.log
python bench.py > gpurun_out/s2_bench_final.json 2> gpurun_out/s2_bench_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/s2_bench_ref.json 2> gpurun_out/s2_bench_ref.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/s2_bench_ref.json
