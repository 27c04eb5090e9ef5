mkdir -p gpurun_out
N=${1:-2}
(timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_parity.py 2>&1 | grep "dist_parity" ) > gpurun_out/dist_parity_n$N.log
cat gpurun_out/dist_parity_n$N.log
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --profile-stages > gpurun_out/bench_n$N.json) 2> gpurun_out/bench_n$N.err
grep -v "^W\|^\[W\|warn" gpurun_out/bench_n$N.err | tail -18; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_n$N.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['check'])"
for mdl in ialspp safer2pp; do
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 --model $mdl --profile-stages > gpurun_out/bench_${mdl}_n$N.json) 2> gpurun_out/bench_${mdl}_n$N.err
grep -v "^W\|^\[W\|warn" gpurun_out/bench_${mdl}_n$N.err | tail -14; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_${mdl}_n$N.json')); print('$mdl', d['ms_per_step'], d['e2e']['ms_per_step'], d['check'])"
done
