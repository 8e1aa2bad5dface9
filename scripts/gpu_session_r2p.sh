#!/bin/bash
# ncu source-line attribution of the shared-memory kernel after the sampled-level search
O=gpurun_out
python scripts/profile_target2.py small > $O/r2p_plain_small.log 2>&1 || { echo "plain failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:pf_small_kernel -s 2 -c 1 -o /tmp/r2p_small python scripts/profile_target2.py small > $O/r2p_ncu_small.log 2>&1
python scripts/ncu_metrics.py /tmp/r2p_small.ncu-rep $O/r2p_small_summary.json > $O/r2p_small_metrics.txt 2>&1
python scripts/ncu_lines.py /tmp/r2p_small.ncu-rep _ZN3sgm15pf_small_kernelIfNS_8SvmPriorELi512ELi2ELb1EEEvNS_5KArgsE 45 > $O/r2p_small_lines.txt 2>&1
ncu -i /tmp/r2p_small.ncu-rep --page source --csv --print-source sass > $O/r2p_small_sass.csv 2>/dev/null
ls -la $O/r2p_*; head -70 $O/r2p_small_lines.txt
