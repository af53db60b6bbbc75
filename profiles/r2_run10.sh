set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2j; mkdir -p $O
timeout 300 python bench.py --steps 50 --warmup 5 --no-model --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
ls -la $O
