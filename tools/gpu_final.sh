#!/bin/bash
# Round-end verification on one B200: gpurun -- "bash tools/gpu_final.sh" -> pytest -m gpu, smoke(), default bench.
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -q -m gpu --timeout 400 2>&1 | tail -25) > gpurun_out/final_tests.log
tail -6 gpurun_out/final_tests.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4) > gpurun_out/final_smoke.log; cat gpurun_out/final_smoke.log
(timeout 400 python bench.py > gpurun_out/final_bench.json) 2> gpurun_out/final_bench.err
tail -3 gpurun_out/final_bench.err; python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['gpu_launches'], d['clocks'])"
