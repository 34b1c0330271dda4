#!/bin/bash
# ncu of kernel 1 on the two small configurations AFTER the item-order changes
O=gpurun_out
python profiles/time_asm.py c3 > $O/plain_c3_final.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 2 -c 1 -o $O/r2zz_asm_c3 -f python profiles/time_asm.py c3 > /dev/null 2>&1
python profiles/time_asm.py c1 > $O/plain_c1_final.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:assemble_kernel -s 2 -c 1 -o $O/r2zz_asm_c1 -f python profiles/time_asm.py c1 > /dev/null 2>&1
cat $O/plain_c3_final.log $O/plain_c1_final.log | cut -c1-160
ls -la $O/r2zz_*.ncu-rep
