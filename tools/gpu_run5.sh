#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
VARIANTS="32:2.0:8192,2:2.0:8192,3:2.0:8192,4:2.0:8192,32:2.0:16384,3:2.0:16384" timeout 600 python tools/sweep_variants.py > $O/r2_variants_v5.jsonl 2> $O/r2_variants_v5.err; echo "rc=$?" >> $O/r2_variants_v5.err
timeout 600 python -m pytest tests/test_gpu_fused_c2.py tests/test_gpu_fused.py -x -q -m gpu > $O/r2_t5.log 2>&1; echo "rc=$?" >> $O/r2_t5.log
cat $O/r2_variants_v5.jsonl | cut -c1-200
tail -3 $O/r2_variants_v5.err $O/r2_t5.log
