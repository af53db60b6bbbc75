cd $GRAFT_REPO_ROOT
O=gpurun_out/r4q; mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,launch__grid_size,launch__block_size
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?" >> $O/bench_ref.err
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?" >> $O/bench_n1.err
python profiles/run_stage_once.py > $O/wtlayer.plain 2>&1 && timeout 300 ncu --metrics $M --clock-control none --csv -k regex:'k_tcgemm|cst' -s 40 -c 40 --log-file $O/wtlayer.csv python profiles/run_stage_once.py > $O/wtlayer.ncu.log 2>&1
