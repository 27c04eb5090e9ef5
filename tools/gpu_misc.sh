mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -x -k "msd_configuration or block_solvers or fixture_train" 2>&1 | tail -5) > gpurun_out/m_tests.log
tail -3 gpurun_out/m_tests.log
run() { # name, args...
  name=$1; shift
  (timeout 500 python bench.py --no-cpu-baseline --profile-stages "$@" > gpurun_out/m_$name.json) 2> gpurun_out/m_$name.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/m_$name.json')); print('$name', round(d['ms_per_step'],2), 'ms/epoch e2e', round(d['e2e']['ms_per_step'],2), {k: round(v,2) for k,v in d['roofline']['stage_ms'].items() if v > 0.3})
except Exception as e:
    print('$name FAILED', e); print(open('gpurun_out/m_$name.err').read()[-600:])
PY
}
run ialspp_d64 --model ialspp --dim 64 --steps 2 --warmup 1
run ialspp_d128 --model ialspp --dim 128 --steps 2 --warmup 1
run ialspp_d256 --model ialspp --dim 256 --steps 2 --warmup 1
run safer2pp_d256 --model safer2pp --dim 256 --steps 2 --warmup 1
run ials_d256 --model ials --dim 256 --steps 3 --warmup 2
run msd_ials_d512 --shape msd --model ials --dim 512 --steps 1 --warmup 1
(FRX_TC_DEBUG=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --dim 128 > /dev/null) 2> gpurun_out/m_dbg128.err
grep "frx tc\|frx wb" gpurun_out/m_dbg128.err | tail -5
