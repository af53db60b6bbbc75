set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2h; mkdir -p $O
timeout 600 python profiles/du_bias_probe.py > $O/du_bias.jsonl 2> $O/du_bias.err
timeout 900 python -m pytest tests/test_fullmodel_gpu.py -m gpu -q > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log
ls -la $O
