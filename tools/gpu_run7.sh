#!/bin/bash
# K1 TMA-staged gather: parity suite + A/B bench + ncu; full GPU suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -q -m gpu > $O/r2_t7.log 2>&1; echo "rc=$?" >> $O/r2_t7.log
timeout 400 python tests/perf/bench_kernels.py feat > $O/r2_kernels_feat.jsonl 2> $O/r2_kernels_feat.err; echo "rc=$?" >> $O/r2_kernels_feat.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:feat_forward -s 4 -c 2 -o $O/r2_prof_feat_v1 python tests/perf/profile_hbm.py > $O/r2_ncu_feat.log 2>&1
tail -15 $O/r2_t7.log
cat $O/r2_kernels_feat.jsonl | cut -c1-300
tail -3 $O/r2_kernels_feat.err $O/r2_ncu_feat.log
