cd $GRAFT_REPO_ROOT
O=gpurun_out/r3k; mkdir -p $O
timeout 600 python -m pytest tests/test_block_gpu.py -m gpu -q -k capturable > $O/pytest_graph.log 2>&1; echo "rc=$?" >> $O/pytest_graph.log
