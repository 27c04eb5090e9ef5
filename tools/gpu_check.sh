mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 120 -x -s -k "dual_form or tensor_core_row or fixture_train" 2>&1 | tail -40) > gpurun_out/a_tests.log
tail -30 gpurun_out/a_tests.log
if grep -q "passed" gpurun_out/a_tests.log && ! grep -q "failed" gpurun_out/a_tests.log; then
(FRX_TC_DEBUG=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/a_dbg.json) 2> gpurun_out/a_dbg.err
(timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-stages > gpurun_out/a_bench.json) 2> gpurun_out/a_bench.err
grep "frx wb" gpurun_out/a_dbg.err | tail -2; cat gpurun_out/a_bench.err | tail -12; python -c "
import json; d=json.load(open('gpurun_out/a_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['check'])"
fi
