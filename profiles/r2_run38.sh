cd $GRAFT_REPO_ROOT
O=gpurun_out/r3l; mkdir -p $O
timeout 900 python -m pytest tests/test_wtconv_gpu.py -m gpu -q > $O/pytest_wt.log 2>&1; echo "rc=$?" >> $O/pytest_wt.log
