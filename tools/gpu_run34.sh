#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_properties.py tests/test_gpu_fused.py tests/test_gpu_fused_c2.py -q -m gpu > $O/r2_t34.log 2>&1; echo "rc=$?" >> $O/r2_t34.log
tail -6 $O/r2_t34.log | cut -c1-250
PROF_USERS=400000 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:finalize -c 6 --csv --log-file $O/r2_fin_time_v34.csv python tools/profile_embed.py --batch 16384 --steps 6 > $O/r2_ncu_fin34.log 2>&1
grep finalize $O/r2_fin_time_v34.csv | awk -F'","' '{print $NF}' | tr '\n' ' '; echo
VARIANTS="213:2.0:16384:-12,213:2.0:32768:-12,213:2.0:16384:-12" timeout 500 python tools/sweep_variants.py > $O/r2_variants_v34.jsonl 2>/dev/null; cut -c1-230 $O/r2_variants_v34.jsonl
timeout 600 python bench.py --steps 50 --no-cpu-baseline > $O/r2_bench_v34.json 2> $O/r2_bench_v34.err; echo "rc=$?" >> $O/r2_bench_v34.err
python -c "
import json; d=json.load(open('$O/r2_bench_v34.json')); print(d['value'], d['e2e'], d['parity'], d['fallback_rows'])"
