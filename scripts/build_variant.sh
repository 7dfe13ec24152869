#!/bin/bash
# scripts/build_variant.sh NAME "EXTRA NVCC FLAGS": an alternative build of the library into cugraph_b200/lib_NAME (A/B experiments)
set -e
N=$1; shift
cd "$(dirname "$0")/.."
O=cugraph_b200/csrc/_obj_$N; L=cugraph_b200/lib_$N
mkdir -p $O $L
FLAGS="-O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC,-fvisibility=hidden -I include -I cugraph_b200/csrc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a"
pids=()
for f in cugraph_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  /usr/local/cuda/bin/nvcc $FLAGS $@ -Xptxas=-v -c $f -o $O/$b.o > $O/$b.log 2>&1 &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $L/libcugraph_c.so $O/*.o -ccbin /usr/bin/g++ -Xcompiler -fPIC
/usr/local/cuda/bin/nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I include scripts/cbench.cu -o $L/cbench -L $L -l:libcugraph_c.so -Xlinker -rpath -Xlinker '$ORIGIN' -ccbin /usr/bin/g++
grep -A2 "k_sweepIfLb0" $O/pagerank.log | grep -v "^--" | head -8
