#!/bin/bash
# ncu of the SHIPPED kernel 1 (lean cexp, mirror-paired item order) -- the capture traffic.json refers to
set -x
O=gpurun_out
python profiles/time_asm.py c1 8192 > $O/plain_asm8192_final.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 2 -c 1 -o $O/r2z_asm_n8192 -f python profiles/time_asm.py c1 8192 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 2 -c 1 -o $O/r2z_asm_c1 -f python profiles/time_asm.py c1 > /dev/null 2>&1
python profiles/time_dense.py sym 8192 > $O/plain_dense8192_final.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:shard_update_kernel -s 40 -c 1 -o $O/r2z_dense_update_n8192 -f python profiles/time_dense.py sym 8192 > /dev/null 2>&1
cat $O/plain_asm8192_final.log $O/plain_dense8192_final.log | cut -c1-120
ls -la $O/r2z_*.ncu-rep
