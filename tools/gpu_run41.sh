#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_lightgcn_sharded.py -q -m gpu > $O/r2_t41.log 2>&1; echo "rc=$?" >> $O/r2_t41.log
tail -30 $O/r2_t41.log | cut -c1-300
