#!/bin/bash
O=gpurun_out
( time timeout 1500 python bench.py > $O/r2x_bench_n1.log 2> $O/r2x_bench_n1.err ) 2> $O/r2x_bench_time.txt; echo "bench rc=$?"; tail -3 $O/r2x_bench_time.txt
tail -1 $O/r2x_bench_n1.log | cut -c1-1500
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2x_bench_ref.log 2> $O/r2x_bench_ref.err ) 2> $O/r2x_ref_time.txt; echo "ref rc=$?"; tail -3 $O/r2x_ref_time.txt
tail -1 $O/r2x_bench_ref.log | cut -c1-800
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
