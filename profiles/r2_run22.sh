cd $GRAFT_REPO_ROOT
O=gpurun_out/r2v; mkdir -p $O
timeout 600 python -m pytest tests/test_block_gpu.py tests/test_attention_gpu.py -m gpu -q > $O/pytest_block.log 2>&1; echo "rc=$?" >> $O/pytest_block.log
timeout 200 python profiles/block_breakdown.py 32 32 128 32 > $O/block_d32.txt 2>&1
timeout 900 python -m pytest tests/test_fullmodel_gpu.py -m gpu -q > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
timeout 300 python bench_model.py train --steps 5 > $O/train_dropin.json 2> $O/train_dropin.err
