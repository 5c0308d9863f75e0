#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
VARIANTS="32:2.0:8192:0:20000,32:2.0:8192:0:2000,32:2.0:8192:0:200,32:2.0:8192:0:0,32:2.0:8192:1:20000,32:2.0:8192:1:200,32:2.0:8192:1:0,4:2.0:8192:0:0,3:2.0:8192:0:0" timeout 600 python tools/sweep_variants.py > $O/r2_variants_v8.jsonl 2> $O/r2_variants_v8.err; echo "rc=$?" >> $O/r2_variants_v8.err
cat $O/r2_variants_v8.jsonl | cut -c1-180
tail -3 $O/r2_variants_v8.err
