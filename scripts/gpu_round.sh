# Round evidence: default bench (both arms), ncu launch list of the same command, ncu --set full of the layer kernels.
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -c 300 gpurun_out/bench_full.err
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
$B > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu_launch.log 2>&1
B2="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-graph"
$B2 > gpurun_out/plain_full.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:spmm_|premask' -s 15 -c 5 -o gpurun_out/prof_r1_final $B2 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
