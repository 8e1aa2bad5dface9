#!/bin/bash
# A/B builds of libsgmpf.so: scripts/build_variant.sh <name> "<extra nvcc flags>"  ->  <pkg>/libsgmpf_<name>.so
# (select at run time with SGM_LIB_PATH=<that file>; the variants are git-ignored like the main library).
cd "$(dirname "$0")/.." || exit 1
PKG="stochastic-gradient-mcmc-for-non-linear-state-models---mth422_b200"
name="$1"; extra="$2"
obj="$PKG/build_$name"; mkdir -p "$obj"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v $extra"
pids=()
for u in sgmpf sgmpf_f32 sgmpf_f64 sgmpf_cl32 sgmpf_cl64; do
  /usr/local/cuda/bin/nvcc $FLAGS -c -o "$obj/$u.o" "$PKG/csrc/$u.cu" > "$obj/$u.log" 2>&1 &
  pids+=($!)
done
rc=0; for p in "${pids[@]}"; do wait "$p" || rc=1; done
[ $rc -ne 0 ] && { grep -h error "$obj"/*.log | head; exit 1; }
/usr/local/cuda/bin/nvcc -shared -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -o "$PKG/libsgmpf_$name.so" "$obj"/*.o
echo "built $PKG/libsgmpf_$name.so"
