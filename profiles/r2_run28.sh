cd $GRAFT_REPO_ROOT
O=gpurun_out/r3b; mkdir -p $O
timeout 600 python -m pytest tests/test_block_gpu.py -m gpu -q > $O/pytest_block.log 2>&1; echo "rc=$?" >> $O/pytest_block.log
timeout 200 python profiles/block_breakdown.py 32 32 128 32 > $O/block_d32.txt 2>&1
