#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_two_tower_train.py tests/test_gpu_feat_models.py tests/test_gpu_movielens_c1.py tests/test_gpu_multi_sparse.py tests/test_gpu_dynamic.py tests/test_gpu_dropin.py tests/test_gpu_fm_train.py tests/test_gpu_deepfm_train.py -q -m gpu > $O/r2_t19.log 2>&1; echo "rc=$?" >> $O/r2_t19.log
tail -30 $O/r2_t19.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:feat_forward -s 2 -c 1 -o $O/r2_prof_feat_v19 python tests/perf/profile_hbm.py > $O/r2_ncu_feat19.log 2>&1
tail -2 $O/r2_ncu_feat19.log
VARIANTS="213:2.67:16384,213:2.67:16384:-8,213:2.0:16384:-8,213:1.6:16384:-8,213:2.0:16384:-4,213:1.3:16384:-12" timeout 500 python tools/sweep_variants.py > $O/r2_variants_v19.jsonl 2> $O/r2_variants_v19.err; echo "rc=$?" >> $O/r2_variants_v19.err
cut -c1-230 $O/r2_variants_v19.jsonl; tail -2 $O/r2_variants_v19.err
