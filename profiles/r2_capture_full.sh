# the D = 32 mixer chain with --set full (the .ncu-rep, ~22 MB, fits the 64 MiB return limit) + the launch list of the bench command
cd $GRAFT_REPO_ROOT
O=gpurun_out/r2ncu; mkdir -p $O
python profiles/run_mixer_once.py > $O/mixer_d32.plain 2>&1 && ncu --set full --clock-control none --import-source on -s 9 -c 9 -o $O/mixer_d32 -f python profiles/run_mixer_once.py > $O/mixer_d32.ncu.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu --no-model > $O/bench_short.json 2> $O/bench_short.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-model > $O/launches.ncu.log 2>&1
du -sh $O
