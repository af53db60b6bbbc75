set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2k; mkdir -p $O
N=${1:-2}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-model > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "rc=$?" >> $O/bench_n$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 200 --warmup 5 --no-model > $O/bench_n${N}_200.json 2> $O/bench_n${N}_200.err; echo "rc=$?" >> $O/bench_n${N}_200.err
ls -la $O
