#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
VARIANTS="113:2.0:8192,213:2.0:8192,110:2.0:8192,112:2.0:8192,4113:2.0:8192,4213:2.0:8192,213:1.6:8192,212:1.6:8192,4213:1.6:8192,213:2.0:16384,4213:2.0:16384" timeout 500 python tools/sweep_variants.py > $O/r2_variants_v11.jsonl 2> $O/r2_variants_v11.err; echo "rc=$?" >> $O/r2_variants_v11.err
cat $O/r2_variants_v11.jsonl | cut -c1-200
tail -3 $O/r2_variants_v11.err
timeout 600 python -m pytest tests/test_gpu_fused_c2.py tests/test_gpu_fused.py -x -q -m gpu > $O/r2_t11.log 2>&1; echo "rc=$?" >> $O/r2_t11.log
tail -3 $O/r2_t11.log
