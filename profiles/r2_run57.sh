cd $GRAFT_REPO_ROOT
O=gpurun_out/r4r; mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,launch__grid_size,launch__block_size
python profiles/run_stage_once.py > $O/wtlayer.plain 2>&1 && timeout 400 ncu --metrics $M --clock-control none --csv -s 64 -c 64 --log-file $O/wtlayer.csv python profiles/run_stage_once.py > $O/wtlayer.ncu.log 2>&1
