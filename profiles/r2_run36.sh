cd $GRAFT_REPO_ROOT
O=gpurun_out/r3j; mkdir -p $O
timeout 300 python profiles/module_times.py 256 64 2 dropin infer > $O/module_times_infer.json 2> $O/module_times_infer.err
