cd $GRAFT_REPO_ROOT
O=gpurun_out/r3m; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?" >> $O/bench_ref.err
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
