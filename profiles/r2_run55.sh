cd $GRAFT_REPO_ROOT
O=gpurun_out/r4p; mkdir -p $O
timeout 200 python -m pytest tests/test_tcgemm_gpu.py -q -x --timeout 120 > $O/pytest_gemm.log 2>&1; echo "rc=$?" >> $O/pytest_gemm.log
if grep -q "rc=0" $O/pytest_gemm.log; then
timeout 600 python -m pytest tests/test_convstage_gpu.py tests/test_block_gpu.py tests/test_mixer_gpu.py tests/test_attention_gpu.py -q --timeout 300 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python bench_model.py train --steps 8 --warmup 3 > $O/train.json 2> $O/train.err
timeout 300 python bench_model.py infer --steps 5 --warmup 2 > $O/infer.json 2> $O/infer.err
fi
