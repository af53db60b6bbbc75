cd $GRAFT_REPO_ROOT
O=gpurun_out/r4s; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 170 python profiles/sweep.py > $O/sweep.log 2>&1; echo "rc=$?" >> $O/sweep.log; cp gpurun_out/r02_sweep.json $O/r02_sweep.json
