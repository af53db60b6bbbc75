cd $GRAFT_REPO_ROOT
O=gpurun_out/r4b; mkdir -p $O
timeout 300 python -m pytest tests/test_convstage_gpu.py -q --timeout 300 > $O/pytest_convstage.log 2>&1; echo "rc=$?" >> $O/pytest_convstage.log
timeout 300 python profiles/stage_breakdown.py wtlayer 128 32 > $O/stage_wtlayer_dec6.txt 2>&1
timeout 300 python profiles/stage_breakdown.py wtlayer 64 32 32 64 2 0 > $O/stage_wtlayer_enc2.txt 2>&1
timeout 300 python profiles/stage_breakdown.py outproj 128 32 > $O/stage_outproj.txt 2>&1
timeout 300 python profiles/stage_breakdown.py patchembed 128 32 > $O/stage_patchembed.txt 2>&1
timeout 300 python profiles/fullmodel_errs.py 1 > $O/errs.json 2> $O/errs.err
timeout 300 python profiles/module_times.py 128 32 2 dropin > $O/module_times.json 2> $O/module_times.err
