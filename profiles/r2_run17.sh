cd $GRAFT_REPO_ROOT
O=gpurun_out/r2q; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
timeout 300 python profiles/module_times.py 128 32 2 dropin > $O/module_times.json 2> $O/module_times.err
