set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2i; mkdir -p $O
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "rc=$?" >> $O/bench_n$N.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-graph-nccl --no-model > $O/bench_n${N}_side.json 2> $O/bench_n${N}_side.err; echo "rc=$?" >> $O/bench_n${N}_side.err
ls -la $O
