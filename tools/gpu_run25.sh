#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_deepfm_train.py tests/test_gpu_fm_train.py tests/test_gpu_two_tower_train.py tests/test_gpu_youtube_ranking_train.py tests/test_gpu_lightgcn.py -q -m gpu -x > $O/r2_t25.log 2>&1; echo "rc=$?" >> $O/r2_t25.log
tail -30 $O/r2_t25.log
timeout 400 python bench.py --config c3 > $O/r2_bench_c3_v25.json 2> $O/r2_bench_c3_v25.err; echo "rc=$?" >> $O/r2_bench_c3_v25.err
cut -c1-400 $O/r2_bench_c3_v25.json; tail -c 600 $O/r2_bench_c3_v25.json; tail -3 $O/r2_bench_c3_v25.err
