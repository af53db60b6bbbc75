cd $GRAFT_REPO_ROOT
O=gpurun_out/r4w; mkdir -p $O
timeout 75 python -m pytest tests -m gpu -q -x > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
