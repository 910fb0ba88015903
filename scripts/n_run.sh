# usage: . scripts/n_run.sh; nrun TAG NGPU [ENV=..]... -- [bench args]   (multi-GPU A/B helper)
nrun() {
  TAG=$1; N=$2; shift 2
  ENVS=""; while [ "$1" != "--" ] && [ $# -gt 0 ]; do ENVS="$ENVS $1"; shift; done; shift
  PORT=$((29500 + RANDOM % 400))
  env $ENVS timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 20 --warmup 5 --extras off "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
    print("$TAG", "N", j["n_gpus"], "ms/step", round(j["ms_per_step"],4), "exchange", j["run"]["exchange"], "barrier:", j["run"].get("arrival_barrier"), "e2e", (j.get("e2e") or {}).get("ms_per_step"))
except Exception as e:
    print("$TAG failed", e); print(open("gpurun_out/bench_$TAG.err").read()[-1500:])
PY
}
