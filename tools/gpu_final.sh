#!/bin/bash
# final validation: what the driver runs at round end (GPU tests, smoke, reference arm, own arm)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r2_final_tests.log 2>&1; echo "rc=$?" >> $O/r2_final_tests.log
tail -4 $O/r2_final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_final_smoke.log 2>&1; echo "rc=$?" >> $O/r2_final_smoke.log
tail -3 $O/r2_final_smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > $O/r2_final_bench_ref.json 2> $O/r2_final_bench_ref.err; echo "rc=$?" >> $O/r2_final_bench_ref.err
cut -c1-700 $O/r2_final_bench_ref.json; tail -2 $O/r2_final_bench_ref.err
timeout 600 python bench.py > $O/r2_final_bench.json 2> $O/r2_final_bench.err; echo "rc=$?" >> $O/r2_final_bench.err
cut -c1-300 $O/r2_final_bench.json; tail -2 $O/r2_final_bench.err
