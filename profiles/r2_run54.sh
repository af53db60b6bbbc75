cd $GRAFT_REPO_ROOT
O=gpurun_out/r4o; mkdir -p $O
timeout 300 python -m pytest tests/test_convstage_gpu.py tests/test_block_gpu.py tests/test_tcgemm_gpu.py -q --timeout 300 > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 300 python bench_model.py infer --steps 5 --warmup 2 > $O/infer.json 2> $O/infer.err
