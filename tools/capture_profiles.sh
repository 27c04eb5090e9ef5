#!/bin/bash
# Profile capture, run on the GPU box through gpurun (ONE ncu invocation per call, after the same
# command has exited 0 without ncu).  Usage:
#   gpurun -- 'bash tools/capture_profiles.sh launches'   -> gpurun_out/launches.csv
#   gpurun -- 'bash tools/capture_profiles.sh full'       -> gpurun_out/row_solve_tc.ncu-rep (+ csv)
#   gpurun -- 'bash tools/capture_profiles.sh wb'         -> gpurun_out/row_solve_wb.ncu-rep (+ csv)
set -e
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline ${BENCH_ARGS}"
$CMD > gpurun_out/plain_$1.log 2> gpurun_out/plain_$1.err
case "$1" in
  launches)
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
        $CMD > gpurun_out/ncu_launches.log 2>&1 ;;
  full)
    # second epoch's launches of the direct row kernel (pieces, long rows, ordinary rows; both half-steps)
    ncu --set full --clock-control none --import-source on -k regex:row_solve_tc -s 6 -c 6 \
        -o gpurun_out/row_solve_tc -f $CMD > gpurun_out/ncu_full.log 2>&1
    ncu -i gpurun_out/row_solve_tc.ncu-rep --page raw --csv > gpurun_out/row_solve_tc_raw.csv 2>/dev/null ;;
  wb)
    ncu --set full --clock-control none --import-source on -k regex:"row_solve_wb|sym_tridiag|rows_gemm" -s 4 -c 5 \
        -o gpurun_out/row_solve_wb -f $CMD > gpurun_out/ncu_wb.log 2>&1
    ncu -i gpurun_out/row_solve_wb.ncu-rep --page raw --csv > gpurun_out/row_solve_wb_raw.csv 2>/dev/null ;;
  others)
    ncu --set full --clock-control none -k regex:"user_resid|gramian_tc|quadform|user_loss_finish|xi_newton" -s 6 -c 7 \
        -o gpurun_out/stage_kernels -f $CMD > gpurun_out/ncu_others.log 2>&1
    ncu -i gpurun_out/stage_kernels.ncu-rep --page raw --csv > gpurun_out/stage_kernels_raw.csv 2>/dev/null ;;
esac
