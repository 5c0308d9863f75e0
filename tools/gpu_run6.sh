#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
VARIANTS="32:2.0:8192:0,32:2.0:8192:1,32:2.0:8192:2,32:2.0:8192:3,4:2.0:8192:1,4:2.0:8192:2,4:2.0:8192:3,3:2.0:8192:2" timeout 600 python tools/sweep_variants.py > $O/r2_variants_v6.jsonl 2> $O/r2_variants_v6.err; echo "rc=$?" >> $O/r2_variants_v6.err
cat $O/r2_variants_v6.jsonl | cut -c1-160
tail -3 $O/r2_variants_v6.err
