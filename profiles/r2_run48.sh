cd $GRAFT_REPO_ROOT
O=gpurun_out/r4i; mkdir -p $O
timeout 300 python profiles/bridge_probe.py > $O/bridge_probe.txt 2>&1
