# bench line at N GPUs (N = $1), as the driver launches it
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-roofline --no-cpu-baseline > gpurun_out/s2_bench_${N}gpu.json 2> gpurun_out/s2_bench_${N}gpu.err; echo "bench rc=$?"
tail -1 gpurun_out/s2_bench_${N}gpu.json | cut -c1-220
