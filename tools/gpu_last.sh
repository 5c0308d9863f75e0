#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $O/r2_last_ref.json 2> $O/r2_last_ref.err; echo "rc=$?" >> $O/r2_last_ref.err
cut -c1-200 $O/r2_last_ref.json; tail -1 $O/r2_last_ref.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/r2_last_bench.json 2> $O/r2_last_bench.err; echo "rc=$?" >> $O/r2_last_bench.err
cut -c1-200 $O/r2_last_bench.json; tail -1 $O/r2_last_bench.err
