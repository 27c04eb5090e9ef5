mkdir -p gpurun_out
N=4
(timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --profile-stages > gpurun_out/bench_n$N.json) 2> gpurun_out/bench_n$N.err
grep "^{" gpurun_out/bench_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['stage_ms'])"
