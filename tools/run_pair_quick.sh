# pair kernel: parity against the 16-lane kernel + kernel times + quick bench (run under gpurun)
python tools/pair_check.py > gpurun_out/r03_pair_check.json 2> gpurun_out/pair_check.err; echo pair_check rc=$?
python bench.py --steps 30 --warmup 5 --skip-fits --skip-cpu > gpurun_out/r03_bench_quick.json 2> gpurun_out/r03_bench_quick.err; echo bench rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/r03_pair_check.json'))
print(d['ok'])
a,b=d['kernel_ms_16lane'],d['kernel_ms_pair']
for k in a: print("%-20s K2 16-lane %.4f  pair %.4f" % (k, a[k][1], b[k][1]))
bq=json.loads(open('gpurun_out/r03_bench_quick.json').read().strip().splitlines()[-1])
print({k:bq[k] for k in ('value','ms_per_step')}, bq['e2e']['value'], {k:v['ms'] for k,v in bq['roofline']['kernels'].items()})
PY
