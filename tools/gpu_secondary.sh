#!/bin/bash
# Secondary configurations on one B200, run through gpurun:
#   gpurun -- 'bash tools/gpu_secondary.sh'   -> gpurun_out/sec_*.json (one bench line each)
mkdir -p gpurun_out
run() { # name, bench.py args...
  name=$1; shift
  (timeout 600 python bench.py --no-cpu-baseline --profile-stages "$@" > gpurun_out/sec_$name.json) 2> gpurun_out/sec_$name.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/sec_$name.json'))
    print('$name', round(d['ms_per_step'], 2), 'ms/epoch, e2e', round(d['e2e']['ms_per_step'], 2),
          {k: round(v, 2) for k, v in d['roofline']['stage_ms'].items() if v > 0.3})
except Exception as e:
    print('$name FAILED', e)
PY
}
run safer2_d128 --dim 128
run ials_d256 --model ials
run cvar_mf_d256 --model cvar_mf --steps 3 --warmup 2
run ialspp_d64 --model ialspp --dim 64 --steps 2 --warmup 1
run ialspp_d128 --model ialspp --dim 128 --steps 2 --warmup 1
run ialspp_d256 --model ialspp --dim 256 --steps 2 --warmup 1
run safer2pp_d256 --model safer2pp --dim 256 --steps 2 --warmup 1
run msd_ials_d512 --shape msd --model ials --dim 512 --steps 1 --warmup 1
run msd_erm_mf_d512 --shape msd --model erm_mf --dim 512 --steps 1 --warmup 1
