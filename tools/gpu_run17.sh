#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 tools/ubench/gather > $O/r2_ubench_gather.jsonl 2> $O/r2_ubench_gather.err; echo "rc=$?" >> $O/r2_ubench_gather.err
cat $O/r2_ubench_gather.jsonl
timeout 300 python tools/lightgcn_shard_probe.py > $O/r2_lg_probe.jsonl 2> $O/r2_lg_probe.err
cat $O/r2_lg_probe.jsonl | cut -c1-1200; tail -3 $O/r2_lg_probe.err
timeout 400 python bench.py --config c5 > $O/r2_bench_c5.json 2> $O/r2_bench_c5.err; echo "rc=$?" >> $O/r2_bench_c5.err
cut -c1-900 $O/r2_bench_c5.json; tail -2 $O/r2_bench_c5.err
# one full capture of the shipped sweep (PRE + MAIN of the second 16 384-user launch)
PROF_USERS=400000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 2 -c 2 -o $O/r2_prof_sweep_v17 python tools/profile_embed.py --batch 16384 --steps 3 > $O/r2_ncu_sweep17.log 2>&1
tail -2 $O/r2_ncu_sweep17.log
PROF_USERS=400000 timeout 600 ncu --set full --clock-control none --import-source on -k regex:finalize -s 1 -c 1 -o $O/r2_prof_finalize_v17 python tools/profile_embed.py --batch 16384 --steps 3 > $O/r2_ncu_fin17.log 2>&1
tail -2 $O/r2_ncu_fin17.log
# launch list of the bench command (serialised, cold-cache: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_v17.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2_launches_bench.log 2>&1
tail -1 $O/r2_launches_bench.log | cut -c1-300
# shared-memory race check + memcheck of the fused path at a small shape
PROF_USERS=20000 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/profile_embed.py --items 60000 --batch 512 --steps 1 > $O/r2_racecheck.log 2>&1; echo "rc=$?" >> $O/r2_racecheck.log
tail -6 $O/r2_racecheck.log
PROF_USERS=20000 timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/profile_embed.py --items 60000 --batch 512 --steps 1 > $O/r2_memcheck.log 2>&1; echo "rc=$?" >> $O/r2_memcheck.log
tail -6 $O/r2_memcheck.log
