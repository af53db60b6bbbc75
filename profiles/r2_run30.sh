cd $GRAFT_REPO_ROOT
O=gpurun_out/r3d; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
timeout 200 python profiles/block_breakdown.py 32 32 128 32 > $O/block_d32.txt 2>&1
timeout 900 python profiles/sweep.py > $O/sweep.log 2>&1; cp gpurun_out/r02_sweep.json $O/r02_sweep.json
