#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
VARIANTS="123:2.0:8192:1,129:2.0:8192:1,128:2.0:8192:1,113:2.0:8192:1,119:2.0:8192:1,118:2.0:8192:1,113:2.0:8192:0" timeout 400 python tools/sweep_variants.py > $O/r2_variants_v10.jsonl 2> $O/r2_variants_v10.err; echo "rc=$?" >> $O/r2_variants_v10.err
cat $O/r2_variants_v10.jsonl | cut -c1-160
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"sweep_kernel|guess_kernel|finalize_kernel|prep_users" -s 8 -c 16 --csv --log-file $O/r2_launches_v2.csv python tools/profile_embed.py --steps 4 --batch 8192 > $O/r2_ncu_l2.log 2>&1
tail -3 $O/r2_variants_v10.err
grep -c sweep $O/r2_launches_v2.csv
