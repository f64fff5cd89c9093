mkdir -p gpurun_out
python bench.py --batch-sweep --no-roofline --no-cpu-baseline > gpurun_out/s3_sweep.json 2> gpurun_out/s3_sweep.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/s3_sweep.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
print(json.dumps(d.get('batch_sweep'), indent=1))
"
tail -3 gpurun_out/s3_sweep.err
timeout 300 python tools/exp_tiled.py 10 c4x10 > gpurun_out/s3_x10.txt 2>&1; echo "x10 rc=$?"
cat gpurun_out/s3_x10.txt | tail -5
