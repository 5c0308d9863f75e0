#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_deepfm_train.py tests/test_gpu_fm_train.py tests/test_gpu_two_tower_train.py tests/test_gpu_youtube_ranking_train.py tests/test_gpu_lightgcn.py tests/test_gpu_feat_models.py -q -m gpu -x > $O/r2_t27.log 2>&1; echo "rc=$?" >> $O/r2_t27.log
tail -8 $O/r2_t27.log
timeout 400 python bench.py --config c3 > $O/r2_bench_c3_v27.json 2> $O/r2_bench_c3_v27.err; echo "rc=$?" >> $O/r2_bench_c3_v27.err
tail -c 500 $O/r2_bench_c3_v27.json; tail -3 $O/r2_bench_c3_v27.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_train_v27.csv python tools/profile_train_step.py > $O/r2_launches_train.log 2>&1
tail -2 $O/r2_launches_train.log
timeout 300 python tests/perf/bench_kernels.py train > $O/r2_kernels_train_v27.jsonl 2> $O/r2_kernels_train_v27.err
cut -c1-300 $O/r2_kernels_train_v27.jsonl
