#!/bin/bash
# round-2 GPU call 1: parity of the new K4 on the speculative path, bench + variants, ncu
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/r2_clocks1.csv &
SMI=$!
timeout 900 python -m pytest tests/test_gpu_fused_c2.py tests/test_gpu_fused.py tests/test_gpu_ranking.py -x -q -m gpu -s > $O/r2_t1.log 2>&1; echo "rc=$?" >> $O/r2_t1.log
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2_bench_a.json 2> $O/r2_bench_a.err; echo "rc=$?" >> $O/r2_bench_a.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --epi-warps 3 > $O/r2_bench_w3.json 2> $O/r2_bench_w3.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --pre-coef 2.0 > $O/r2_bench_c20.json 2> $O/r2_bench_c20.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --pre-coef 1.6 > $O/r2_bench_c16.json 2> $O/r2_bench_c16.err
timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --batch 8192 > $O/r2_bench_b8k.json 2> $O/r2_bench_b8k.err
timeout 400 python -m pytest tests/test_gpu_dropin.py -x -q -m gpu > $O/r2_t2.log 2>&1; echo "rc=$?" >> $O/r2_t2.log
timeout 600 python -m pytest tests -x -q -m gpu > $O/r2_t3.log 2>&1; echo "rc=$?" >> $O/r2_t3.log
kill $SMI
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2_launches_v1.csv python tools/profile_embed.py --steps 3 > $O/r2_ncu_l1.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 3 -c 2 -o $O/r2_prof_sweep_v1 python tools/profile_embed.py --steps 3 > $O/r2_ncu_s1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:finalize_kernel -s 1 -c 1 -o $O/r2_prof_fin_v1 python tools/profile_embed.py --steps 2 > $O/r2_ncu_f1.log 2>&1
tail -3 $O/r2_t1.log $O/r2_t2.log $O/r2_t3.log
cat $O/r2_bench_a.json | cut -c1-3000
