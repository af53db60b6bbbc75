cd $GRAFT_REPO_ROOT
O=gpurun_out/r3a; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
