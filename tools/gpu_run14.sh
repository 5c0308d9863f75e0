#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused_c2.py tests/test_gpu_fused.py tests/test_gpu_ranking.py -x -q -m gpu > $O/r2_t14.log 2>&1; echo "rc=$?" >> $O/r2_t14.log
tail -3 $O/r2_t14.log
VARIANTS="215:2.0:8192,115:2.0:8192,213:2.0:8192,215:1.6:8192,215:2.0:16384,215:1.6:16384" timeout 500 python tools/sweep_variants.py > $O/r2_variants_v14.jsonl 2> $O/r2_variants_v14.err; echo "rc=$?" >> $O/r2_variants_v14.err
cat $O/r2_variants_v14.jsonl | cut -c1-200
tail -3 $O/r2_variants_v14.err
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2_bench_b.json 2> $O/r2_bench_b.err; echo "rc=$?" >> $O/r2_bench_b.err
cat $O/r2_bench_b.json | cut -c1-2500
