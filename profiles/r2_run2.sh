set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2b; mkdir -p $O
timeout 600 python -m pytest tests/test_tcgemm_gpu.py -m gpu -q -x > $O/pytest_gemm.log 2>&1; echo "rc=$?" >> $O/pytest_gemm.log
timeout 900 python -m pytest tests/test_mixer_gpu.py -m gpu -q > $O/pytest_mixer.log 2>&1; echo "rc=$?" >> $O/pytest_mixer.log
timeout 900 python -m pytest tests/test_fullmodel_gpu.py -m gpu -q > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
timeout 600 python bench_model.py train --batch 32 --steps 5 --warmup 3 > $O/train_dropin.json 2> $O/train_dropin.err
timeout 600 python bench_model.py breakdown --batch 32 > $O/breakdown_dropin.json 2> $O/breakdown_dropin.err
timeout 600 python profiles/sweep.py > $O/sweep.json 2> $O/sweep.err
ls -la $O
