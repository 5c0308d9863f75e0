#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_c2.py -q -m gpu > $O/r2_t39.log 2>&1; echo "rc=$?" >> $O/r2_t39.log
tail -6 $O/r2_t39.log | cut -c1-250
VARIANTS="213:2.0:16384:-12,216:2.0:16384:-12,213:2.0:16384:-12,216:2.0:16384:-12,216:2.0:8192:-12,213:2.0:8192:-12" timeout 600 python tools/sweep_variants.py > $O/r2_variants_v39.jsonl 2>/dev/null; cut -c1-215 $O/r2_variants_v39.jsonl
