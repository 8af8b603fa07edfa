cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out/r2r
timeout 1500 python tools/configs_bench.py > gpurun_out/r2r/configs.jsonl 2> gpurun_out/r2r/configs.err; echo rc=$?; tail -c 800 gpurun_out/r2r/configs.err
python - <<'P'
import json
for l in open('gpurun_out/r2r/configs.jsonl'):
    d=json.loads(l); print(d['config'][:44], '|', (d.get('content') or '')[:24], '|', round(d['ms'],2), 'ok', d['ok'], 'oracle', d.get('oracle',{}).get('ok'), d.get('stage_ms_last_call'), (d.get('roofline_stage') or {}).get('frac'))
P
