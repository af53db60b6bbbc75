cd $GRAFT_REPO_ROOT
O=gpurun_out/r4v; mkdir -p $O
timeout 60 python -m pytest tests/test_convstage_gpu.py -q -x --timeout 50 -k "container" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
