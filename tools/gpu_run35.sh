#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:linear_tf32x3 -c 1 -o $O/r2_prof_lin_din_v35 python tools/profile_din_all_items.py > $O/r2_ncu_lin35.log 2>&1
tail -2 $O/r2_ncu_lin35.log
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:din_attention_hoisted -c 1 -o $O/r2_prof_att_din_v35 python tools/profile_din_all_items.py > $O/r2_ncu_att35.log 2>&1
tail -2 $O/r2_ncu_att35.log
