#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_deepfm_train.py tests/test_gpu_fm_train.py tests/test_gpu_two_tower_train.py tests/test_gpu_youtube_ranking_train.py tests/test_gpu_din_train.py tests/test_gpu_lightgcn.py -q -m gpu > $O/r2_t38.log 2>&1; echo "rc=$?" >> $O/r2_t38.log
tail -30 $O/r2_t38.log | cut -c1-300
