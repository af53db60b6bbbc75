set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2l; mkdir -p $O
timeout 300 python -m pytest tests/test_tcgemm_gpu.py -m gpu -q -x > $O/pytest_gemm.log 2>&1; echo "rc=$?" >> $O/pytest_gemm.log
timeout 600 python -m pytest tests/test_mixer_gpu.py -m gpu -q > $O/pytest_mixer.log 2>&1; echo "rc=$?" >> $O/pytest_mixer.log
for cfg in "128 16 128 16" "1024 16 8 32" "512 16 16 32" "128 16 16 32"; do
  timeout 200 python profiles/generic_path_breakdown.py $cfg >> $O/wide_breakdown.txt 2>> $O/wide_breakdown.err
done
ls -la $O
