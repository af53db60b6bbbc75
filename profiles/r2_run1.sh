set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2a; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_fullmodel_gpu.py > $O/pytest_old.log 2>&1; echo "rc=$?" >> $O/pytest_old.log
timeout 900 python -m pytest tests/test_fullmodel_gpu.py -m gpu -q > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
timeout 600 python profiles/fullmodel_probe.py 128 2 > $O/probe.json 2> $O/probe.err
timeout 600 python bench_model.py train --batch 32 --steps 5 --warmup 3 > $O/train_dropin.json 2> $O/train_dropin.err
timeout 600 python bench_model.py train --batch 32 --steps 5 --warmup 3 --variant reference > $O/train_ref.json 2> $O/train_ref.err
timeout 600 python bench_model.py breakdown --batch 32 > $O/breakdown_dropin.json 2> $O/breakdown_dropin.err
timeout 600 python bench_model.py breakdown --batch 32 --variant reference > $O/breakdown_ref.json 2> $O/breakdown_ref.err
timeout 600 python bench_model.py infer --batch 64 --img 256 --steps 3 --warmup 1 > $O/infer_dropin.json 2> $O/infer_dropin.err
timeout 600 python bench_model.py infer --batch 64 --img 256 --steps 3 --warmup 1 --variant reference > $O/infer_ref.json 2> $O/infer_ref.err
timeout 600 python bench.py --steps 50 --warmup 5 > $O/bench.json 2> $O/bench.err
ls -la $O
