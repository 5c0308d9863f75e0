#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
VARIANTS="214:2.0:8192,215:2.0:8192,115:2.0:8192,215:1.6:8192,215:2.67:8192,215:2.0:16384,214:2.0:16384" timeout 500 python tools/sweep_variants.py > $O/r2_variants_v13.jsonl 2> $O/r2_variants_v13.err; echo "rc=$?" >> $O/r2_variants_v13.err
cat $O/r2_variants_v13.jsonl | cut -c1-200
tail -3 $O/r2_variants_v13.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 1 -o $O/r2_prof_sweep_v13_e5 python tools/profile_embed.py --steps 3 --batch 8192 --pre-coef 2.0 --epi-warps 215 > $O/r2_ncu_s13.log 2>&1
tail -2 $O/r2_ncu_s13.log
