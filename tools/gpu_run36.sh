#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_feat_models.py tests/test_gpu_linear_tc.py tests/test_gpu_movielens_c1.py tests/test_gpu_dynamic.py -q -m gpu > $O/r2_t36.log 2>&1; echo "rc=$?" >> $O/r2_t36.log
tail -6 $O/r2_t36.log | cut -c1-250
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_din_v36.csv python tools/profile_din_all_items.py > $O/r2_launches_din.log 2>&1
tail -1 $O/r2_launches_din.log
timeout 400 python tests/perf/bench_kernels.py seq > $O/r2_kernels_seq_v36.jsonl 2> $O/r2_kernels_seq_v36.err
cut -c1-260 $O/r2_kernels_seq_v36.jsonl
