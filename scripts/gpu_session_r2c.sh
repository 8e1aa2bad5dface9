#!/bin/bash
O=gpurun_out
for w in small smallbatch f64; do
  python scripts/profile_target2.py $w > $O/r2c_plain_$w.log 2>&1 || { echo "plain run $w failed"; tail -5 $O/r2c_plain_$w.log; continue; }
  k=pf_small_kernel; [ $w = f64 ] && k=pf_step_kernel
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $O/r2c_$w python scripts/profile_target2.py $w > $O/r2c_ncu_$w.log 2>&1
  tail -2 $O/r2c_ncu_$w.log
done
ls -la $O/*.ncu-rep
