#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_linear_tc.py tests/test_gpu_deepfm_train.py tests/test_gpu_fm_train.py tests/test_gpu_two_tower_train.py tests/test_gpu_youtube_ranking_train.py tests/test_gpu_feat_models.py -q -m gpu > $O/r2_t33.log 2>&1; echo "rc=$?" >> $O/r2_t33.log
tail -25 $O/r2_t33.log | cut -c1-250
timeout 400 python bench.py --config c3 > $O/r2_bench_c3_v33.json 2> $O/r2_bench_c3_v33.err; echo "rc=$?" >> $O/r2_bench_c3_v33.err
tail -c 420 $O/r2_bench_c3_v33.json; tail -2 $O/r2_bench_c3_v33.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_train_v33.csv python tools/profile_train_step.py > $O/r2_launches_train.log 2>&1
tail -1 $O/r2_launches_train.log
