# usage: bash scripts/gpu_multi.sh N   -- N-GPU torchrun bench (with and without the output all-gather)
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1500 gpurun_out/bench_n$N.json; tail -c 400 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 5 --no-allgather --no-e2e > gpurun_out/bench_n${N}_nogather.json 2> gpurun_out/bench_n${N}_nogather.err
tail -c 700 gpurun_out/bench_n${N}_nogather.json
