#!/bin/bash
# Round-2 ncu captures (1 GPU).  Launch lists are cold-cache and serialised: compare shares, not absolutes.
set -x
O=gpurun_out
python profiles/prof_c1.py c3 1 > $O/plain_c3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches_c3.csv python profiles/prof_c1.py c3 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches_c1.csv python profiles/prof_c1.py c1 1 > /dev/null 2>&1
python profiles/time_asm.py c1 8192 > $O/plain_asm8192.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 2 -c 1 -o $O/r2_asm_n8192 -f python profiles/time_asm.py c1 8192 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 2 -c 1 -o $O/r2_asm_c1 -f python profiles/time_asm.py c1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 2 -c 1 -o $O/r2_asm_c3 -f python profiles/time_asm.py c3 > /dev/null 2>&1
python profiles/time_dense.py sym 8192 > $O/plain_dense8192.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:shard_update_kernel -s 34 -c 1 -o $O/r2_dense_update_n8192 -f python profiles/time_dense.py sym 8192 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:panel_sym_kernel -s 300 -c 1 -o $O/r2_dense_panel_n8192 -f python profiles/time_dense.py sym 8192 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:ptrace_kernel -s 1 -c 1 -o $O/r2_dense_ptrace_n8192 -f python profiles/time_dense.py sym 8192 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 1500 --csv --log-file $O/r2_launches_dense8192.csv python profiles/time_dense.py sym 8192 > /dev/null 2>&1
python bench.py --quick --steps 2 --warmup 3 > $O/plain_bench_quick.json 2> $O/plain_bench_quick.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file $O/r2_launches_bench.csv python bench.py --quick --steps 2 --warmup 3 > /dev/null 2>&1
ls -la $O/r2_* | head -20
cat $O/plain_asm8192.log $O/plain_dense8192.log
