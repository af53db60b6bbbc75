cd $GRAFT_REPO_ROOT
O=gpurun_out/r4c; mkdir -p $O
timeout 300 python -m pytest tests/test_convstage_gpu.py -q --timeout 300 > $O/pytest_convstage.log 2>&1; echo "rc=$?" >> $O/pytest_convstage.log
timeout 300 python profiles/stage_breakdown.py wtlayer 128 32 > $O/stage_wtlayer_dec6.txt 2>&1
timeout 300 python profiles/stage_breakdown.py outproj 128 32 > $O/stage_outproj.txt 2>&1
timeout 300 python profiles/stage_breakdown.py patchembed 128 32 > $O/stage_patchembed.txt 2>&1
timeout 300 python bench_model.py train --steps 8 --warmup 3 > $O/train.json 2> $O/train.err
timeout 300 python bench_model.py infer --steps 5 --warmup 2 > $O/infer.json 2> $O/infer.err
timeout 600 python -m pytest tests/test_fullmodel_gpu.py -q --timeout 500 > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
