#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_feat_models.py tests/test_gpu_linear_tc.py tests/test_gpu_dynamic.py -q -m gpu -x > $O/r2_t29.log 2>&1; echo "rc=$?" >> $O/r2_t29.log
tail -12 $O/r2_t29.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_din_v29.csv python tools/profile_din_all_items.py > $O/r2_launches_din.log 2>&1
tail -2 $O/r2_launches_din.log
timeout 400 python tests/perf/bench_kernels.py seq > $O/r2_kernels_seq_v29.jsonl 2> $O/r2_kernels_seq_v29.err
cut -c1-300 $O/r2_kernels_seq_v29.jsonl
