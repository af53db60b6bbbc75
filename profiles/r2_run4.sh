set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2d; mkdir -p $O
timeout 900 python -m pytest tests/test_mixer_gpu.py -m gpu -q > $O/pytest_mixer.log 2>&1; echo "rc=$?" >> $O/pytest_mixer.log
timeout 900 python -m pytest tests/test_fullmodel_gpu.py -m gpu -q > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
for cfg in "128 16 128 16" "1024 16 8 32" "512 16 16 32" "32 128 128 16"; do
  timeout 300 python profiles/generic_path_breakdown.py $cfg >> $O/wide_breakdown.txt 2>> $O/wide_breakdown.err
done
timeout 600 python bench_model.py train --batch 32 --steps 5 --warmup 3 > $O/train_dropin.json 2> $O/train_dropin.err
ls -la $O
