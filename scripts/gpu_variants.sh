python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t4.log; tail -3 gpurun_out/t4.log
B="python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline"
for v in b200 mb4 mb2; do
  SAGNN_B200_LIB=$PWD/sa-gnn_b200/lib/libsagnn_$v.so $B > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - <<PY
import json
j=json.loads(open("gpurun_out/bench_$v.json").read().strip().splitlines()[-1])
print("$v", "ms/step", round(j["ms_per_step"],4), "fwd", round(j["roofline"]["fwd_ms"],4), "bwd", round(j["roofline"]["bwd_ms"],4), "frac", round(j["roofline"]["frac"],3))
PY
done
