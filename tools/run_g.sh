cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out/r2t
for sk in 1 0; do
TDOA_FFT_SKEW=$sk timeout 600 python bench.py --steps 10 --warmup 3 --no-sharded --no-oracle-check --no-cpu-restatement > gpurun_out/r2t/bench_skew$sk.json 2> gpurun_out/r2t/bench_skew$sk.err; echo skew=$sk rc=$?; tail -c 300 gpurun_out/r2t/bench_skew$sk.err
python - <<P
import json
d=json.loads(open('gpurun_out/r2t/bench_skew$sk.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['serial_ms_per_step'], d['parity_check']['lags_match_injected_delays'])
for r in d['roofline_kernels']: print(r['kernel'], round(r['frac'],3), round(r['kernel_ms_per_launch'],3))
P
done
TDOA_FFT_SKEW=1 timeout 900 python -m pytest tests -m gpu -x -q -k "fft or golden or xcorr or full" > gpurun_out/r2t/pytest.txt 2>&1; tail -n 3 gpurun_out/r2t/pytest.txt
