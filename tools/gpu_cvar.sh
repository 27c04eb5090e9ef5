mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "tensor_core or stage_parity or epoch_parity or train_to_host or checkpoint" 2>&1 | tail -30) > gpurun_out/v_tests.log
tail -30 gpurun_out/v_tests.log
(timeout 300 python bench.py --no-cpu-baseline --profile-stages --model cvar_mf --steps 3 --warmup 2 > gpurun_out/v_cvar.json) 2> gpurun_out/v_cvar.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/v_cvar.json')); print('cvar_mf d256', round(d['ms_per_step'],2), 'ms/epoch e2e', round(d['e2e']['ms_per_step'],2), {k: round(v,2) for k,v in d['roofline']['stage_ms'].items() if v > 0.05}, d['check'])
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/v_cvar.err').read()[-800:])
PY
