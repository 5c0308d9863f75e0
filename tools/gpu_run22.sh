#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_k1_variants.py tests/test_gpu_feat_models.py tests/test_gpu_two_tower_train.py tests/test_gpu_deepfm_train.py tests/test_gpu_fm_train.py -q -m gpu -x > $O/r2_t22.log 2>&1; echo "rc=$?" >> $O/r2_t22.log
tail -30 $O/r2_t22.log
timeout 400 python tests/perf/bench_kernels.py feat > $O/r2_kernels_feat_v6.jsonl 2> $O/r2_kernels_feat_v6.err; echo "rc=$?" >> $O/r2_kernels_feat_v6.err
cut -c1-330 $O/r2_kernels_feat_v6.jsonl | head -12; tail -3 $O/r2_kernels_feat_v6.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:feat_forward -s 2 -c 1 -o $O/r2_prof_feat_v22 python tests/perf/profile_hbm.py > $O/r2_ncu_feat22.log 2>&1
tail -2 $O/r2_ncu_feat22.log
