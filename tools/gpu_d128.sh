mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "tensor_core or dual_form or goldens or smoke" 2>&1 | tail -8) > gpurun_out/w_tests.log
tail -4 gpurun_out/w_tests.log
(FRX_TC_DEBUG=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --dim 128 > /dev/null) 2> gpurun_out/w_dbg128.err
grep "frx tc" gpurun_out/w_dbg128.err | tail -4
(timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-stages --dim 128 > gpurun_out/w_bench128.json) 2> gpurun_out/w_bench128.err
grep "step_" gpurun_out/w_bench128.err; python -c "
import json; d=json.load(open('gpurun_out/w_bench128.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['check'])"
