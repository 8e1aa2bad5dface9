#!/bin/bash
# final evidence of round 2: full GPU suite, f32 model table, bench launch list + full ncu capture of the step kernel
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r3e_tests.log 2>&1; echo "pytest rc=$?" >> $O/r3e_tests.log
tail -4 $O/r3e_tests.log
timeout 600 python scripts/probe_models.py --dtype f32 --json $O/r3e_models_f32.json 2>&1 | cut -c1-112 | tail -9
CMD="python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline"
$CMD > $O/r3e_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file /tmp/r3e_launches.csv $CMD > $O/r3e_ncu_bench.log 2>&1
python scripts/ncu_launch_list.py /tmp/r3e_launches.csv $O/r3e_launch_list_bench.json "$CMD" > $O/r3e_launch_list.txt 2>&1; head -8 $O/r3e_launch_list.txt
ncu --set full --clock-control none --import-source on -k regex:pf_step_kernel -s 40 -c 1 -o /tmp/r3e_step $CMD > $O/r3e_ncu_step.log 2>&1
python scripts/ncu_metrics.py /tmp/r3e_step.ncu-rep $O/r3e_step_summary.json > $O/r3e_step_metrics.txt 2>&1
python scripts/ncu_lines.py /tmp/r3e_step.ncu-rep _ZN3sgm14pf_step_kernelIfNS_8SvmPriorELb1ELi1ELb0EEEvNS_5KArgsEi 30 > $O/r3e_step_lines.txt 2>&1
head -14 $O/r3e_step_metrics.txt
