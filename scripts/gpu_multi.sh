# usage: bash scripts/gpu_multi.sh N [modes]  -- N-GPU torchrun bench: default (zero-copy row-block all-to-all), all-gather, compute only
N=${1:-2}
P=29517
for X in ${2:-auto allgather none}; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 5 --exchange $X --no-e2e > gpurun_out/bench_n${N}_$X.json 2> gpurun_out/bench_n${N}_$X.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_n${N}_$X.json").read().strip().splitlines()[-1])
    print("N=$N exchange=$X value %.3e ms/step %.4f"%(j["value"], j["ms_per_step"]))
except Exception as e:
    print("N=$N $X FAILED", e); print(open("gpurun_out/bench_n${N}_$X.err").read()[-600:])
PY
done
