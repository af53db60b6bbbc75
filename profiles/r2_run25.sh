cd $GRAFT_REPO_ROOT
O=gpurun_out/r2y; mkdir -p $O
timeout 600 python -m pytest tests/test_wtconv_gpu.py -m gpu -q > $O/pytest_wt.log 2>&1; echo "rc=$?" >> $O/pytest_wt.log
timeout 200 python profiles/wtconv_launches.py > $O/wt_launches.txt 2>&1
