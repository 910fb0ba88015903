# usage: bash scripts/gpu_profile.sh <tag> [extra bench args]   (one ncu run per gpurun call)
TAG=$1; shift
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-graph $@"
$B > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:spmm_|premask' -s 15 -c 5 -o gpurun_out/prof_$TAG $B > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
