cd $GRAFT_REPO_ROOT
O=gpurun_out/r2o; mkdir -p $O
timeout 300 python -m pytest tests/test_tcgemm_gpu.py -m gpu -q > $O/pytest_gemm.log 2>&1; echo "rc=$?" >> $O/pytest_gemm.log
timeout 300 python profiles/gemm_knockout.py > $O/knockout.txt 2> $O/knockout.err
