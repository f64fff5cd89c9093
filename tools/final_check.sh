# End-of-round check on one B200: the whole GPU suite (no -x: every failure is listed), smoke(), the full bench line.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s3_t_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s3_t_final.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/s3_t_final.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/s3_smoke.log
python bench.py --batch-sweep > gpurun_out/s3_bench_final.json 2> gpurun_out/s3_bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/s3_bench_final.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['launches_per_step'])
for k in ('config4', 'config5', 'config5_reference_shape', 'train_loop_resident', 'batch_sweep'):
    print(k, json.dumps(d.get(k))[:300])
print('roofline', d['roofline']['frac'], d['roofline']['bwd']['frac'])
PY
tail -2 gpurun_out/s3_bench_final.err
