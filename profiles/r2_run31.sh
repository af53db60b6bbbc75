cd $GRAFT_REPO_ROOT
O=gpurun_out/r3e; mkdir -p $O
timeout 600 python -m pytest tests/test_wtconv_gpu.py -m gpu -q > $O/pytest_wt.log 2>&1; echo "rc=$?" >> $O/pytest_wt.log
timeout 200 python profiles/wtconv_launches.py > $O/wt_launches.txt 2>&1
ADN_TIME=1 ADN_GRID=32 ADN_B=4096 timeout 200 python profiles/run_wtconv_once.py > $O/wt_32.txt 2>&1
ADN_TIME=1 ADN_GRID=64 ADN_B=1024 timeout 200 python profiles/run_wtconv_once.py > $O/wt_64.txt 2>&1
