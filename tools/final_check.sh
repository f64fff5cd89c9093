# End-of-round check on one B200: the whole GPU suite, smoke(), the full bench line (with the batch sweep).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s3_t_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_t_final.log
tail -2 gpurun_out/s3_t_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/s3_smoke.log
python bench.py --batch-sweep > gpurun_out/s3_bench_final.json 2> gpurun_out/s3_bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/s3_bench_final.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['launches_per_step'])
for k in ('config4', 'config5', 'config5_reference_shape', 'train_loop_e2e', 'train_loop_resident', 'cpu_baseline', 'gpu_eager_baseline', 'batch_sweep', 'dp_check'):
    print(k, json.dumps(d.get(k))[:400])
print('roofline', d['roofline']['frac'], d['roofline']['bwd']['frac'], d['roofline'].get('traffic_source'))
PY
tail -3 gpurun_out/s3_bench_final.err
