#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_youtube_ranking_train.py tests/test_gpu_two_tower_train.py tests/test_gpu_lightgcn.py tests/test_gpu_lightgcn_sharded.py tests/test_gpu_dropin.py tests/test_gpu_feat_models.py -q -m gpu > $O/r2_t24.log 2>&1; echo "rc=$?" >> $O/r2_t24.log
tail -25 $O/r2_t24.log
timeout 300 python tools/lightgcn_shard_probe.py > $O/r2_lg_probe_v24.jsonl 2> $O/r2_lg_probe_v24.err
cat $O/r2_lg_probe_v24.jsonl | cut -c1-400; tail -3 $O/r2_lg_probe_v24.err
timeout 400 python bench.py --config c5 > $O/r2_bench_c5_v24.json 2> $O/r2_bench_c5_v24.err; echo "rc=$?" >> $O/r2_bench_c5_v24.err
cut -c1-600 $O/r2_bench_c5_v24.json
