cd $GRAFT_REPO_ROOT
O=gpurun_out/r4e; mkdir -p $O
timeout 300 python bench_model.py train --steps 8 --warmup 3 > $O/train.json 2> $O/train.err
timeout 600 python -m pytest tests/test_fullmodel_gpu.py -q --timeout 500 -k "graph_mode" > $O/pytest_graph.log 2>&1; echo "rc=$?" >> $O/pytest_graph.log
