mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 120 -x -k "tensor_core or bench_configuration or dual_form or goldens or epoch_parity or long_rows" 2>&1 | tail -5) > gpurun_out/x_tests.log
tail -3 gpurun_out/x_tests.log
for v in "" _hionly; do
  export FRECSYS_B200_LIB=$PWD/safer2-recommender_b200/libfrecsys_b200$v.so
  (FRX_TC_DEBUG=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null) 2> gpurun_out/x_dbg$v.err
  echo "== variant '$v'"; grep "frx tc" gpurun_out/x_dbg$v.err | tail -4
  (timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --profile-stages > gpurun_out/x_bench$v.json) 2> gpurun_out/x_bench$v.err
  grep "step_" gpurun_out/x_bench$v.err
done
