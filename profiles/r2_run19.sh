cd $GRAFT_REPO_ROOT
O=gpurun_out/r2s; mkdir -p $O
timeout 300 python profiles/fullmodel_errs.py 1 > $O/errs_block1.json 2> $O/errs_block1.err
timeout 300 python profiles/fullmodel_errs.py 0 > $O/errs_block0.json 2> $O/errs_block0.err
