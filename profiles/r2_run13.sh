set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2m; mkdir -p $O
python profiles/generic_path_breakdown.py 128 16 128 16 > $O/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_tcgemm' -s 28 -c 14 -o $O/gemm_d128 python profiles/generic_path_breakdown.py 128 16 128 16 > $O/ncu.log 2>&1
ls -la $O
