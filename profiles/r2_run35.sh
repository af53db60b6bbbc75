cd $GRAFT_REPO_ROOT
O=gpurun_out/r3i; mkdir -p $O
timeout 900 python -m pytest tests/test_fullmodel_gpu.py -m gpu -q > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
timeout 300 python bench_model.py train --steps 5 > $O/train_dropin.json 2> $O/train_dropin.err
timeout 300 python bench_model.py infer --steps 3 > $O/infer_dropin.json 2> $O/infer_dropin.err
timeout 300 python profiles/module_times.py 128 32 2 dropin > $O/module_times.json 2> $O/module_times.err
