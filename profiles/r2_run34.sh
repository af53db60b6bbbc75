cd $GRAFT_REPO_ROOT
O=gpurun_out/r3h; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
ADN_SM_RESERVE=0 timeout 400 $T --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 10 --no-cpu --no-model > $O/n2_reserve0.json 2> $O/n2_reserve0.err
timeout 400 $T --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 10 --no-cpu --no-model > $O/n2_reserve1.json 2> $O/n2_reserve1.err
ADN_SM_RESERVE=2 timeout 400 $T --master-port 29523 bench.py --gpus 2 --steps 20 --warmup 10 --no-cpu --no-model > $O/n2_reserve2.json 2> $O/n2_reserve2.err
ADN_SM_RESERVE=0 timeout 400 $T --master-port 29524 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu --no-model > $O/n2_reserve0_200.json 2> $O/n2_reserve0_200.err
timeout 400 $T --master-port 29525 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu --no-model > $O/n2_reserve1_200.json 2> $O/n2_reserve1_200.err
