#!/bin/bash
# ncu evidence of round 2, summarised ON THE BOX (the reports themselves exceed the 64 MiB that travel back)
O=gpurun_out
prof() {   # name, kernel regex, skip, count, command...
  name=$1; k=$2; s=$3; c=$4; shift 4
  "$@" > $O/r2h_plain_$name.log 2>&1 || { echo "plain run $name failed"; tail -3 $O/r2h_plain_$name.log; return; }
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o /tmp/r2h_$name "$@" > $O/r2h_ncu_$name.log 2>&1
  python scripts/ncu_metrics.py /tmp/r2h_$name.ncu-rep $O/r2h_${name}_summary.json > $O/r2h_${name}_metrics.txt 2>&1
  ncu -i /tmp/r2h_$name.ncu-rep --page raw --csv > $O/r2h_${name}_raw.csv 2>/dev/null
}
prof paris 'paris_(ar|exact|guide)_kernel' 12 3 python scripts/profile_target2.py paris
prof f64 pf_step_kernel 6 1 python scripts/profile_target2.py f64
prof f64v32 pf_step_kernel 6 1 python scripts/profile_target2.py f64v32
prof small pf_small_kernel 2 1 python scripts/profile_target2.py small
prof step pf_step_kernel 40 1 python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline
for n in paris f64 step small; do
  kern=$(python - <<PY
import json
d=json.load(open('gpurun_out/r2h_${n}_summary.json'))
print(d['launches'][0]['Kernel Name'])
PY
)
  echo "$n: $kern"
done
# per-source-line attribution of the dominant kernels (needs the mangled names)
python scripts/ncu_lines.py /tmp/r2h_paris.ncu-rep "$(cuobjdump -elf stochastic-*/libsgmpf.so | grep -o '_ZN3sgm15paris_ar_kernelIfNS_12GarchOptimalEEEvNS_5KArgsEii' | head -1)" 30 > $O/r2h_paris_lines.txt 2>&1
python scripts/ncu_lines.py /tmp/r2h_f64.ncu-rep _ZN3sgm14pf_step_kernelIdNS_8SvmPriorELb1ELi1ELb0EEEvNS_5KArgsEi 30 > $O/r2h_f64_lines.txt 2>&1
python scripts/ncu_lines.py /tmp/r2h_step.ncu-rep _ZN3sgm14pf_step_kernelIfNS_8SvmPriorELb1ELi1ELb0EEEvNS_5KArgsEi 30 > $O/r2h_step_lines.txt 2>&1
python scripts/ncu_lines.py /tmp/r2h_small.ncu-rep _ZN3sgm15pf_small_kernelIfNS_8SvmPriorELi512ELi2ELb1EEEvNS_5KArgsE 30 > $O/r2h_small_lines.txt 2>&1
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/r2h_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/r2h_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/r2h_ncu_bench.log 2>&1
du -sh $O; ls -la $O/r2h_* | head -40
