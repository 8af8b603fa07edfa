cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out/r2s
python bench.py --steps 20 --warmup 3 > gpurun_out/r2s/bench.json 2> gpurun_out/r2s/bench.err; echo bench rc=$?; tail -c 300 gpurun_out/r2s/bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2s/bench_ref.json 2> gpurun_out/r2s/bench_ref.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s/launches.csv python bench.py --steps 2 --warmup 1 --no-sharded --no-oracle-check --no-cpu-restatement > gpurun_out/r2s/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_demod_df|k_fft_tiles|k_boxcar_small|k_corr_candidates" -c 8 -o gpurun_out/r2s/prof python bench.py --steps 1 --warmup 1 --no-sharded --no-oracle-check --no-cpu-restatement > gpurun_out/r2s/ncu_full.log 2>&1
ls -la gpurun_out/r2s
