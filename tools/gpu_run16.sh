#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/lightgcn_shard_probe.py > $O/r2_lg_probe.jsonl 2> $O/r2_lg_probe.err
cat $O/r2_lg_probe.jsonl
timeout 400 python tests/perf/bench_kernels.py feat > $O/r2_kernels_feat_v2.jsonl 2> $O/r2_kernels_feat_v2.err; echo "rc=$?" >> $O/r2_kernels_feat_v2.err
head -5 $O/r2_kernels_feat_v2.jsonl | cut -c1-220
for c in c3 c4 c5 c1; do timeout 400 python bench.py --config $c > $O/r2_bench_$c.json 2> $O/r2_bench_$c.err; echo "rc=$?" >> $O/r2_bench_$c.err; cut -c1-700 $O/r2_bench_$c.json; tail -2 $O/r2_bench_$c.err; done
timeout 600 python -m pytest tests/test_gpu_feat_models.py tests/test_gpu_movielens_c1.py tests/test_gpu_multi_sparse.py tests/test_gpu_dynamic.py tests/test_gpu_fm_train.py tests/test_gpu_deepfm_train.py -q -m gpu > $O/r2_t16.log 2>&1; echo "rc=$?" >> $O/r2_t16.log
tail -4 $O/r2_t16.log
