cd $GRAFT_REPO_ROOT
O=gpurun_out/r4u; mkdir -p $O
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 100 python -m pytest tests/test_tcgemm_gpu.py tests/test_convstage_gpu.py -q -x --timeout 90 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
