set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_final.log 2>&1; tail -3 gpurun_out/t_final.log
python bench.py > gpurun_out/bench_r1v5.json 2> gpurun_out/bench_r1v5.err
for wl in config3 config4 config5 mixed; do python bench.py --workload $wl --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_r1v5_$wl.json 2> gpurun_out/bench_r1v5_$wl.err; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1v5_ref.json 2> gpurun_out/bench_r1v5_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_v5_launches.csv python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_l5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ramp_convert --launch-skip 5 -c 1 -o gpurun_out/prof_r1v5 python bench.py --streams 1024 --seconds 2 --no-e2e --no-cpu-baseline --steps 4 > gpurun_out/ncu_f5.log 2>&1
tail -2 gpurun_out/ncu_f5.log
