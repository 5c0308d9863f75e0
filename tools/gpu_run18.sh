#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_two_tower_train.py tests/test_gpu_feat_models.py tests/test_gpu_movielens_c1.py tests/test_gpu_multi_sparse.py tests/test_gpu_dynamic.py tests/test_gpu_dropin.py tests/test_gpu_fm_train.py tests/test_gpu_deepfm_train.py -q -m gpu -x > $O/r2_t18.log 2>&1; echo "rc=$?" >> $O/r2_t18.log
tail -30 $O/r2_t18.log
timeout 400 python tests/perf/bench_kernels.py feat > $O/r2_kernels_feat_v3.jsonl 2> $O/r2_kernels_feat_v3.err; echo "rc=$?" >> $O/r2_kernels_feat_v3.err
cut -c1-330 $O/r2_kernels_feat_v3.jsonl; tail -3 $O/r2_kernels_feat_v3.err
