cd $GRAFT_REPO_ROOT
O=gpurun_out/r2t; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
timeout 200 python profiles/block_breakdown.py 32 32 128 32 > $O/block_d32.txt 2>&1
timeout 200 python profiles/block_breakdown.py 256 512 8 32 > $O/block_d256.txt 2>&1
