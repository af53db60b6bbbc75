cd $GRAFT_REPO_ROOT
O=gpurun_out/r2n; mkdir -p $O
timeout 300 python profiles/gemm_knockout.py > $O/knockout.txt 2> $O/knockout.err
