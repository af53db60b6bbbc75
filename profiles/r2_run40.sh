cd $GRAFT_REPO_ROOT
O=gpurun_out/r4a; mkdir -p $O
timeout 600 python -m pytest tests/test_convstage_gpu.py -q -x --timeout 300 > $O/pytest_convstage.log 2>&1; echo "rc=$?" >> $O/pytest_convstage.log
timeout 600 python -m pytest tests/test_convstage_gpu.py -q --timeout 300 > $O/pytest_convstage_all.log 2>&1; echo "rc=$?" >> $O/pytest_convstage_all.log
timeout 600 python -m pytest tests/test_fullmodel_gpu.py tests/test_tcgemm_gpu.py -q --timeout 500 > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
