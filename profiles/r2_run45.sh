cd $GRAFT_REPO_ROOT
O=gpurun_out/r4f; mkdir -p $O
timeout 300 python bench_model.py breakdown > $O/breakdown.txt 2> $O/breakdown.err
timeout 600 python -m pytest tests/test_fullmodel_gpu.py -q --timeout 500 -k "graph_mode" > $O/pytest_graph.log 2>&1; echo "rc=$?" >> $O/pytest_graph.log
