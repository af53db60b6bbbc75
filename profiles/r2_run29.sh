cd $GRAFT_REPO_ROOT
O=gpurun_out/r3c; mkdir -p $O
timeout 300 python -m pytest tests/test_tcgemm_gpu.py tests/test_block_gpu.py -m gpu -q > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python profiles/gemm_knockout.py > $O/knockout.txt 2> $O/knockout.err
timeout 200 python profiles/block_breakdown.py 32 32 128 32 > $O/block_d32.txt 2>&1
timeout 200 python profiles/generic_path_breakdown.py 128 16 128 16 > $O/wide_d128.txt 2>&1
