cd $GRAFT_REPO_ROOT
O=gpurun_out/r3g; mkdir -p $O
timeout 300 python -m pytest tests/test_tcgemm_gpu.py -m gpu -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python profiles/gemm_knockout.py > $O/knockout.txt 2> $O/knockout.err
ADN_VARIANT=16 timeout 300 python profiles/gemm_knockout.py > $O/knockout_4st.txt 2>> $O/knockout.err
