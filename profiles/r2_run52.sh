cd $GRAFT_REPO_ROOT
O=gpurun_out/r4m; mkdir -p $O
timeout 300 python profiles/infer_kernels.py > $O/infer_kernels.json 2> $O/infer_kernels.err
