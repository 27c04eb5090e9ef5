mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "msd_configuration or large_dims or use_cg" 2>&1 | tail -8) > gpurun_out/y_tests.log
tail -4 gpurun_out/y_tests.log
if grep -q passed gpurun_out/y_tests.log && ! grep -q failed gpurun_out/y_tests.log; then
for mdl in ials; do
(timeout 400 python bench.py --no-cpu-baseline --profile-stages --shape msd --model $mdl --dim 512 --steps 1 --warmup 1 > gpurun_out/y_msd_$mdl.json) 2> gpurun_out/y_msd_$mdl.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/y_msd_$mdl.json')); print('msd $mdl d512', round(d['ms_per_step'],1), 'ms/epoch e2e', round(d['e2e']['ms_per_step'],1), {k: round(v,1) for k,v in d['roofline']['stage_ms'].items() if v > 1})
except Exception as e:
    print('FAILED', e); print(open('gpurun_out/y_msd_$mdl.err').read()[-600:])
PY
done
fi
