mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "dual_form or fixture_train or bench_configuration or goldens or tensor_core_row" 2>&1 | tail -8) > gpurun_out/z_tests.log
tail -4 gpurun_out/z_tests.log
(FRX_TC_DEBUG=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null) 2> gpurun_out/z_dbg.err
grep "frx wb" gpurun_out/z_dbg.err | tail -1
(timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-stages > gpurun_out/z_bench.json) 2> gpurun_out/z_bench.err
grep "step_" gpurun_out/z_bench.err; python -c "
import json; d=json.load(open('gpurun_out/z_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['check'])"
