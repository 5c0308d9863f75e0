#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_feat_models.py tests/test_gpu_din_train.py tests/test_gpu_dynamic.py tests/test_gpu_movielens_c1.py -q -m gpu > $O/r2_t40.log 2>&1; echo "rc=$?" >> $O/r2_t40.log
tail -25 $O/r2_t40.log | cut -c1-260
timeout 400 python bench.py --config c4 > $O/r2_bench_c4_v40.json 2> $O/r2_bench_c4_v40.err; echo "rc=$?" >> $O/r2_bench_c4_v40.err
cut -c1-300 $O/r2_bench_c4_v40.json; tail -2 $O/r2_bench_c4_v40.err
