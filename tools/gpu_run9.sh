#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
# parity first (small), under a tight timeout: a cluster dead-lock must not hang the box
timeout 300 python -m pytest tests/test_gpu_fused.py -x -q -m gpu > $O/r2_t9a.log 2>&1; echo "rc=$?" >> $O/r2_t9a.log
tail -3 $O/r2_t9a.log
VARIANTS="223:2.0:8192,213:2.0:8192,123:2.0:8192,113:2.0:8192,223:2.0:16384,213:2.0:16384,220:2.0:8192,223:2.0:8192:1" timeout 400 python tools/sweep_variants.py > $O/r2_variants_v9.jsonl 2> $O/r2_variants_v9.err; echo "rc=$?" >> $O/r2_variants_v9.err
cat $O/r2_variants_v9.jsonl | cut -c1-200
tail -3 $O/r2_variants_v9.err
timeout 600 python -m pytest tests/test_gpu_fused_c2.py -x -q -m gpu > $O/r2_t9b.log 2>&1; echo "rc=$?" >> $O/r2_t9b.log
tail -3 $O/r2_t9b.log
