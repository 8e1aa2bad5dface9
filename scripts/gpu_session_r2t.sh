#!/bin/bash
# ncu counters of the FAST step kernel: register staging (default build) against TMA bulk staging (-DSGM_TMA_WINDOW=1)
O=gpurun_out
P=$PWD/stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:pf_step_kernel -s 40 -c 3 --csv --log-file $O/r2t_ncu_base.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/r2t_ncu_base.log 2>&1
SGM_LIB_PATH=$P/libsgmpf_tma.so ncu --metrics $M --clock-control none -k regex:pf_step_kernel -s 40 -c 3 --csv --log-file $O/r2t_ncu_tma.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > $O/r2t_ncu_tma.log 2>&1
python - <<'PY'
import csv, json, collections
out = {}
for name in ("base", "tma"):
    rows = [r for r in csv.reader(open("gpurun_out/r2t_ncu_%s.csv" % name)) if len(r) > 10]
    hdr = rows[0]; iM, iV, iK, iID = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Kernel Name"), hdr.index("ID")
    acc = collections.defaultdict(list)
    for r in rows[1:]:
        try: acc[r[iM]].append(float(r[iV].replace(",", "")))
        except ValueError: pass
    out[name] = {k: sum(v) / len(v) for k, v in acc.items()}
    out[name]["kernel"] = rows[1][iK]
json.dump(out, open("gpurun_out/r2t_tma_ab_ncu.json", "w"), indent=1)
for k in sorted(out["base"]):
    print(k.ljust(80), out["base"][k], out["tma"].get(k))
PY
