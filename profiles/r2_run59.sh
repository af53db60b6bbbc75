cd $GRAFT_REPO_ROOT
O=gpurun_out/r4t; mkdir -p $O
python profiles/run_stage_once.py > $O/wtlayer.plain 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tcgemm -s 11 -c 3 -o $O/conv_gemm -f python profiles/run_stage_once.py > $O/conv_gemm.ncu.log 2>&1
ls -la $O
