cd $GRAFT_REPO_ROOT
O=gpurun_out/r2r; mkdir -p $O
timeout 600 python -m pytest tests/test_block_gpu.py -m gpu -q > $O/pytest_block.log 2>&1; echo "rc=$?" >> $O/pytest_block.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_block_gpu.py > $O/pytest_rest.log 2>&1; echo "rc=$?" >> $O/pytest_rest.log
timeout 300 python profiles/module_times.py 128 32 2 dropin > $O/module_times.json 2> $O/module_times.err
timeout 300 python bench_model.py train --steps 5 > $O/train_dropin.json 2> $O/train_dropin.err
