set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2g; mkdir -p $O
timeout 900 python bench.py --steps 100 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?" >> $O/bench_ref.err
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
ls -la $O
