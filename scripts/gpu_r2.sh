# usage: [TESTSEL="pytest -k expr"] [BARGS="--workload ..."] bash scripts/gpu_r2.sh [lib variants]
# round 2: gpu tests on the default (packet-stream) kernel, then bench default / v8 / lib/libsagnn_<variant>.so
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q ${TESTSEL:+-k "$TESTSEL"} 2>&1 | tail -30 > gpurun_out/test_r2.log; tail -5 gpurun_out/test_r2.log
run() {  # tag, env...
  TAG=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline $BARGS > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    print("$TAG", "ms/step", round(j["ms_per_step"],4), "fwd", round(j["roofline"]["fwd_ms"],4), "bwd", round(j["roofline"]["bwd_ms"],4), "frac", round(j["roofline"]["frac"],3), "step_frac", round(j["roofline"]["step"]["frac"],3))
except Exception as e:
    print("$TAG failed", e); print(open("gpurun_out/bench_$TAG.err").read()[-2000:])
PY
}
run pkt SAGNN_KERNEL=v9
[ -z "$NOV8" ] && run v8 SAGNN_KERNEL=v8
for v in "$@"; do run $v SAGNN_B200_LIB=$PWD/sa-gnn_b200/lib/libsagnn_$v.so; done
