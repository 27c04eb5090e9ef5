mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -q -m gpu --timeout 400 2>&1 | tail -25) > gpurun_out/f_tests.log
tail -12 gpurun_out/f_tests.log
(timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-stages > gpurun_out/f_bench.json) 2> gpurun_out/f_bench.err
cat gpurun_out/f_bench.err | tail -14; python -c "
import json; d=json.load(open('gpurun_out/f_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['check'])"
(timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-stages --dim 128 > gpurun_out/f_bench128.json) 2> gpurun_out/f_bench128.err
cat gpurun_out/f_bench128.err | tail -14; python -c "
import json; d=json.load(open('gpurun_out/f_bench128.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['check'])"
