#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python tools/profile_din_all_items.py > $O/r2_din_plain.log 2>&1; tail -1 $O/r2_din_plain.log
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2_launches_din_v28.csv python tools/profile_din_all_items.py > $O/r2_launches_din.log 2>&1
tail -2 $O/r2_launches_din.log
