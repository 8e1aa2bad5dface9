#!/bin/bash
# 8-GPU box, final library of round 2: bench at 8 / 4 / 2 ranks (torchrun), chains demo at 8 ranks
O=gpurun_out
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) "$@"; }
for n in 8 4 2; do
  timeout 600 bash -c "$(declare -f run); run $n bench.py --gpus $n --steps 5 --warmup 3" > $O/r3c_bench_n$n.json 2> $O/r3c_bench_n$n.err; echo "bench n=$n rc=$?"
done
for N in 1000 10000; do
  it=2000; [ $N = 10000 ] && it=200
  timeout 400 bash -c "$(declare -f run); run 8 scripts/chains_demo.py --chains 64 --N $N --iters $it --host-iters 10 --json $O/r3c_chains_n8_N$N.json" > $O/r3c_chains_n8_N$N.log 2>&1; echo "chains n=8 N=$N rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r3c_bench_n*.json')):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f, d['n_gpus'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['extra']['config5_strong']['ms_per_gradient'], d['extra']['config5_strong']['frac'],
              d['extra']['config4_chains']['N1000']['chain_iterations_per_sec'], d['extra']['config4_chains']['N10000']['chain_iterations_per_sec'], d['f64']['native']['value'])
    except Exception as e:
        print(f, 'ERR', e)
for f in sorted(glob.glob('gpurun_out/r3c_chains_n*.json')):
    d = json.load(open(f)); print(f, d['n_gpus'], d['device_loop_chain_iterations_per_sec'], d['host_ensemble_chain_iterations_per_sec'])
PY
