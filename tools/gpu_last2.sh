#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_feat_models.py -q -m gpu -k "youtube_retrieval or two_tower or wide_deep" > $O/r2_t_last2.log 2>&1; echo "rc=$?" >> $O/r2_t_last2.log
tail -30 $O/r2_t_last2.log | cut -c1-300
