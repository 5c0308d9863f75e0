#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_properties.py tests/test_gpu_fused.py tests/test_gpu_fused_c2.py tests/test_gpu_ranking.py tests/test_gpu_dropin.py -q -m gpu > $O/r2_t31.log 2>&1; echo "rc=$?" >> $O/r2_t31.log
tail -30 $O/r2_t31.log
VARIANTS="213:2.0:16384" timeout 300 python tools/sweep_variants.py > $O/r2_variants_v31.jsonl 2>/dev/null; cut -c1-260 $O/r2_variants_v31.jsonl
VARIANTS="213:2.0:8192" timeout 300 python tools/sweep_variants.py --dim 256 --items 500000 > $O/r2_variants_v31_d256.jsonl 2>/dev/null; cut -c1-400 $O/r2_variants_v31_d256.jsonl
