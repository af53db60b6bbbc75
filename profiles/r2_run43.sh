cd $GRAFT_REPO_ROOT
O=gpurun_out/r4d; mkdir -p $O
timeout 300 python bench_model.py train --steps 8 --warmup 3 > $O/train.json 2> $O/train.err
timeout 300 python bench_model.py infer --steps 5 --warmup 2 > $O/infer.json 2> $O/infer.err
timeout 600 python -m pytest tests/test_fullmodel_gpu.py -q --timeout 500 -k "graph_mode" > $O/pytest_graph.log 2>&1; echo "rc=$?" >> $O/pytest_graph.log
timeout 300 python -m pytest tests/test_convstage_gpu.py -q --timeout 300 -k "capturable" > $O/pytest_cap.log 2>&1; echo "rc=$?" >> $O/pytest_cap.log
ADNM_TRAIN_GRAPH=0 timeout 300 python bench_model.py train --steps 8 --warmup 3 > $O/train_eager.json 2> $O/train_eager.err
