cd $GRAFT_REPO_ROOT
O=gpurun_out/r4l; mkdir -p $O
timeout 300 python -m pytest tests/test_convstage_gpu.py -q --timeout 300 > $O/pytest_convstage.log 2>&1; echo "rc=$?" >> $O/pytest_convstage.log
timeout 300 python profiles/bridge_probe.py > $O/bridge_probe.txt 2>&1
timeout 300 python bench_model.py train --steps 8 --warmup 3 > $O/train.json 2> $O/train.err
timeout 300 python bench_model.py infer --steps 5 --warmup 2 > $O/infer.json 2> $O/infer.err
