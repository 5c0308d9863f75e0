#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $O/r2_t23.log 2>&1; echo "rc=$?" >> $O/r2_t23.log
tail -8 $O/r2_t23.log
timeout 600 python bench.py > $O/r2_bench_v23.json 2> $O/r2_bench_v23.err; echo "rc=$?" >> $O/r2_bench_v23.err
cut -c1-1500 $O/r2_bench_v23.json; tail -2 $O/r2_bench_v23.err
for c in c3 c4; do timeout 400 python bench.py --config $c > $O/r2_bench_${c}_v23.json 2> $O/r2_bench_${c}_v23.err; echo "rc=$?" >> $O/r2_bench_${c}_v23.err; cut -c1-700 $O/r2_bench_${c}_v23.json; done
timeout 400 python tests/perf/bench_kernels.py seq > $O/r2_kernels_seq_v23.jsonl 2> $O/r2_kernels_seq_v23.err
cut -c1-300 $O/r2_kernels_seq_v23.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 3000 --csv --log-file $O/r2_launches_seq_v23.csv python tests/perf/bench_kernels.py seq > $O/r2_launches_seq.log 2>&1
tail -2 $O/r2_launches_seq.log | cut -c1-200
