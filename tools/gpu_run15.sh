#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
VARIANTS="213:2.0:16384,215:2.0:16384,213:2.0:8192,113:2.0:8192,213:1.6:16384,223:2.0:16384" timeout 500 python tools/sweep_variants.py > $O/r2_variants_v15.jsonl 2> $O/r2_variants_v15.err; echo "rc=$?" >> $O/r2_variants_v15.err
cat $O/r2_variants_v15.jsonl | cut -c1-200
timeout 900 python -m pytest tests -q -m gpu > $O/r2_t15.log 2>&1; echo "rc=$?" >> $O/r2_t15.log
tail -12 $O/r2_t15.log
timeout 400 python tests/perf/bench_kernels.py feat > $O/r2_kernels_feat.jsonl 2> $O/r2_kernels_feat.err; echo "rc=$?" >> $O/r2_kernels_feat.err
cat $O/r2_kernels_feat.jsonl | cut -c1-260
tail -3 $O/r2_kernels_feat.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:feat_forward -s 4 -c 2 -o $O/r2_prof_feat_v1 python tests/perf/profile_hbm.py > $O/r2_ncu_feat.log 2>&1
tail -2 $O/r2_ncu_feat.log
