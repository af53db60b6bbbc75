# ncu evidence of round 2 (each capture only after the same command has exited 0 without ncu).  The .ncu-rep files of the
# wide captures exceed gpurun's 64 MiB return limit, so everything but the D = 32 mixer chain is captured as CSV with the
# metric list of profiles/summarize_ncu.py (a handful of passes instead of --set full's ~40) and only CSVs come back.
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2ncu; mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,launch__grid_size,launch__block_size
NCU="ncu --metrics $M --clock-control none --csv"
if [ "$1" != "nofull" ]; then
python profiles/run_mixer_once.py > $O/mixer_d32.plain 2>&1 && ncu --set full --clock-control none --import-source on -s 9 -c 9 -o $O/mixer_d32 -f python profiles/run_mixer_once.py > $O/mixer_d32.ncu.log 2>&1
ncu -i $O/mixer_d32.ncu-rep --page raw --csv > $O/mixer_d32_raw.csv 2>/dev/null
fi
ADN_D=128 python profiles/run_mixer_once.py > $O/mixer_d128.plain 2>&1 && ADN_D=128 $NCU -s 26 -c 26 --log-file $O/mixer_d128.csv python profiles/run_mixer_once.py > $O/mixer_d128.ncu.log 2>&1
python profiles/run_block_once.py > $O/block_d32.plain 2>&1 && $NCU -s 34 -c 34 --log-file $O/block_d32.csv python profiles/run_block_once.py > $O/block_d32.ncu.log 2>&1
python profiles/run_wtconv_once.py > $O/wtconv.plain 2>&1 && $NCU -s 15 -c 15 --log-file $O/wtconv.csv python profiles/run_wtconv_once.py > $O/wtconv.ncu.log 2>&1
python profiles/run_attention_once.py > $O/attention.plain 2>&1 && $NCU -k regex:'sdpa' -s 6 -c 3 --log-file $O/attention.csv python profiles/run_attention_once.py > $O/attention.ncu.log 2>&1
if [ "$1" != "nofull" ]; then
python bench.py --steps 2 --warmup 3 --no-cpu --no-model > $O/bench_short.json 2> $O/bench_short.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-model > $O/launches.ncu.log 2>&1
fi
du -sh $O; ls -la $O
