# Two-GPU check: the data-parallel tests and the bench line at N = 2.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dp.py -m gpu -x -q > gpurun_out/s2_t_dp2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_t_dp2.log
tail -3 gpurun_out/s2_t_dp2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-roofline --no-cpu-baseline > gpurun_out/s2_bench_2gpu.json 2> gpurun_out/s2_bench_2gpu.err; echo "bench rc=$?"
tail -1 gpurun_out/s2_bench_2gpu.json | cut -c1-200
