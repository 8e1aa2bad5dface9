#!/bin/bash
O=gpurun_out
PKG="stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"
timeout 2400 python -m pytest tests -m gpu -q --maxfail=60 -p no:cacheprovider > $O/r2k_tests.log 2>&1; echo "pytest rc=$?" >> $O/r2k_tests.log
tail -4 $O/r2k_tests.log
for v in default parisu2 parisu1; do
  if [ $v = default ]; then python scripts/probe_paris.py $O/r2k_paris_$v.json > $O/r2k_paris_$v.log 2>&1
  else SGM_LIB_PATH=$PWD/$PKG/libsgmpf_$v.so python scripts/probe_paris.py $O/r2k_paris_$v.json > $O/r2k_paris_$v.log 2>&1; fi
  echo "== $v"; cut -c1-200 $O/r2k_paris_$v.log
done
python scripts/profile_target2.py paris > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:paris_ar_kernel -s 4 -c 1 -o /tmp/r2k_ar python scripts/profile_target2.py paris > $O/r2k_ncu_ar.log 2>&1
python scripts/ncu_metrics.py /tmp/r2k_ar.ncu-rep $O/r2k_paris_ar_summary.json > $O/r2k_paris_ar_metrics.txt 2>&1
python scripts/ncu_lines.py /tmp/r2k_ar.ncu-rep _ZN3sgm15paris_ar_kernelIfNS_12GarchOptimalEEEvNS_5KArgsEii 40 > $O/r2k_paris_ar_lines.txt 2>&1
head -45 $O/r2k_paris_ar_lines.txt | cut -c1-180
grep -E "duration|issue_active|long_score|short_score|warps_active|registers|barrier|lts__t_sector_hit|inst_executed.sum" $O/r2k_paris_ar_metrics.txt
