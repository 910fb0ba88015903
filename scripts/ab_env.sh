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
