#!/bin/bash
# final-library ncu captures of the f64 step kernel and the shared-memory kernel + plain smoke of every kernel family
O=gpurun_out
prof() {   # name, kernel regex, skip, count, command...
  name=$1; k=$2; s=$3; c=$4; shift 4
  "$@" > $O/r3f_plain_$name.log 2>&1 || { echo "plain run $name failed"; tail -3 $O/r3f_plain_$name.log; return; }
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o /tmp/r3f_$name "$@" > $O/r3f_ncu_$name.log 2>&1
  python scripts/ncu_metrics.py /tmp/r3f_$name.ncu-rep $O/r3f_${name}_summary.json > $O/r3f_${name}_metrics.txt 2>&1
}
prof f64 pf_step_kernel 6 1 python scripts/profile_target2.py f64
prof small pf_small_kernel 2 1 python scripts/profile_target2.py small
python scripts/ncu_lines.py /tmp/r3f_f64.ncu-rep _ZN3sgm14pf_step_kernelIdNS_8SvmPriorELb1ELi1ELb0EEEvNS_5KArgsEi 30 > $O/r3f_f64_lines.txt 2>&1
python scripts/ncu_lines.py /tmp/r3f_small.ncu-rep _ZN3sgm15pf_small_kernelIfNS_8SvmPriorELi512ELi2ELb1EEEvNS_5KArgsE 30 > $O/r3f_small_lines.txt 2>&1
head -12 $O/r3f_f64_metrics.txt; head -8 $O/r3f_f64_lines.txt | cut -c1-150
head -12 $O/r3f_small_metrics.txt
timeout 300 python scripts/all_kernels_smoke.py > $O/r3f_all_kernels.log 2>&1; echo "all_kernels rc=$?"; tail -3 $O/r3f_all_kernels.log
timeout 300 python scripts/sanitizer_target.py > $O/r3f_sanitizer_plain.log 2>&1; echo "sanitizer target (plain) rc=$?"; tail -2 $O/r3f_sanitizer_plain.log
