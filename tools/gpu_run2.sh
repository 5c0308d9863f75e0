#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/sweep_variants.py > $O/r2_variants_v2.jsonl 2> $O/r2_variants_v2.err; echo "rc=$?" >> $O/r2_variants_v2.err
timeout 900 python -m pytest tests/test_gpu_fused_c2.py tests/test_gpu_fused.py -x -q -m gpu -s > $O/r2_t4.log 2>&1; echo "rc=$?" >> $O/r2_t4.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 7 -c 1 -o $O/r2_prof_sweep_v2 python tools/profile_embed.py --steps 3 --batch 8192 > $O/r2_ncu_s2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 7 -c 1 -o $O/r2_prof_sweep_v2w4 python tools/profile_embed.py --steps 3 --batch 8192 --epi-warps 4 > $O/r2_ncu_s2w4.log 2>&1
cat $O/r2_variants_v2.jsonl | cut -c1-260
tail -3 $O/r2_t4.log
