# full default bench (both arms) + ncu launch list of the same command; used to fill profiles/
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -c 600 gpurun_out/bench_full.err
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
