#!/bin/bash
# tools/build_variant.sh NAME "-DBWTK_RS_MINB32=4 ..." -> gpurun_variants/libbwtk_NAME.so (tuning builds; not shipped)
set -e
cd "$(dirname "$0")/../bwt-algorithm_b200/csrc"
name=$1; defs=$2
mkdir -p /tmp/bwtk_var_$name ../../gpurun_variants
for f in index lcp sa search fmpack kmer scans extend fasta rowchain; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $defs -c $f.cu -o /tmp/bwtk_var_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_variants/libbwtk_$name.so /tmp/bwtk_var_$name/*.o -lcudart
