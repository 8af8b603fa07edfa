cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out/r2o
python -m pytest tests -m gpu -x -q > gpurun_out/r2o/pytest.txt 2>&1; tail -n 5 gpurun_out/r2o/pytest.txt
python bench.py --steps 10 --warmup 3 --no-sharded > gpurun_out/r2o/bench.json 2> gpurun_out/r2o/bench.err; tail -c 600 gpurun_out/r2o/bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2o/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['serial_ms_per_step'])
for r in d['roofline_kernels']: print(r['kernel'], round(r['frac'],3), round(r['kernel_ms_per_launch'],3))
print(d['parity_check'].get('oracle_at_size'))
P
