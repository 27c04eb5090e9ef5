mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:row_solve_generic -c 2 -o gpurun_out/row_generic_d512 -f \
  python -m pytest tests/test_gpu_parity.py -q -m gpu -k "msd_configuration and ials" > gpurun_out/ncu_generic.log 2>&1
ncu -i gpurun_out/row_generic_d512.ncu-rep --page raw --csv > gpurun_out/row_generic_d512_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_generic.log
