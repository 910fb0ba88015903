# usage: bash scripts/gpu_variants.sh v1 v2 ...   -- bench each lib/libsagnn_<v>.so (no ncu)
for v in "$@"; do
  SAGNN_B200_LIB=$PWD/sa-gnn_b200/lib/libsagnn_$v.so python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms/step", round(j["ms_per_step"],4), "fwd", round(j["roofline"]["fwd_ms"],4), "bwd", round(j["roofline"]["bwd_ms"],4), "frac", round(j["roofline"]["frac"],3))
except Exception as e:
    print("$v failed", e); print(open("gpurun_out/bench_$v.err").read()[-1500:])
PY
done
