#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernel_edges.py tests/test_gpu_k1_variants.py -q -m gpu > $O/r2_t32.log 2>&1; echo "rc=$?" >> $O/r2_t32.log
tail -40 $O/r2_t32.log
