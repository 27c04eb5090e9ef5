#!/bin/bash
# Multi-GPU checks on one box: gpurun --gpus N -- "bash tools/gpu_multi.sh N" -> dist_parity log, bench lines
# (point-to-point and broadcast all-gather), one iALS++ line.
mkdir -p gpurun_out
N=${1:-8}
(DIST_PARITY_CASES="safer2:256,safer2:128,ialspp:32,safer2pp:128,cvar_mf:32" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_parity.py 2>&1 | grep "dist_parity" ) > gpurun_out/dist_parity_n$N.log
cat gpurun_out/dist_parity_n$N.log
for ag in p2p bcast; do
(FRX_ALLGATHER=$ag timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --profile-stages > gpurun_out/bench_n${N}_$ag.json) 2> gpurun_out/bench_n${N}_$ag.err
echo "== allgather $ag"; grep -v "^W\|^\[W\|warn\|^\*\|OMP_NUM\|^$" gpurun_out/bench_n${N}_$ag.err | tail -16; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_n${N}_$ag.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['check'])"
done
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 --model ialspp --profile-stages > gpurun_out/bench_ialspp_n$N.json) 2> gpurun_out/bench_ialspp_n$N.err
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_ialspp_n$N.json')); print('ialspp', d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['stage_ms'])"
