#!/bin/bash
# multi-GPU validation: the collective legs of bench.py and the sharded tests (run with gpurun --gpus N)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 30 --warmup 3 > $O/r2_bench_${N}gpu.json 2> $O/r2_bench_${N}gpu.err; echo "rc=$?" >> $O/r2_bench_${N}gpu.err
timeout 600 python -m pytest tests/test_gpu_lightgcn_sharded.py -x -q -m gpu > $O/r2_t_mg${N}.log 2>&1; echo "rc=$?" >> $O/r2_t_mg${N}.log
tail -c 6000 $O/r2_bench_${N}gpu.json
tail -n 5 $O/r2_bench_${N}gpu.err; tail -n 5 $O/r2_t_mg${N}.log
