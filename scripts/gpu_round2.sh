# round-2 evidence run (1 GPU): full default bench + reference arm + other shapes + ncu launch list + ncu --set full
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; tail -c 400 gpurun_out/r2_bench_full.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
for w in amazon-book ml10m scaled-s10; do python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err; done
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_r2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/plain_r2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"spmm_pkt|premask" -s 10 -c 5 -o gpurun_out/prof_r2_final python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log | cut -c1-200
