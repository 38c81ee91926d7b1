#!/bin/bash
# Build a variant of libfununifrac.so with other compile-time constants, next to the product library:
#   tools/build_variant.sh kc32 -DFUF_KC=32 -DFUF_STAGES=2     ->  fununifrac_b200/csrc/libfununifrac_kc32.so
# Select it for one run with FUF_LIB=$PWD/fununifrac_b200/csrc/libfununifrac_kc32.so (A/B measurements only).
set -e
name=$1; shift
cd "$(dirname "$0")/../fununifrac_b200/csrc"
mkdir -p _build_$name
for f in tree pushup pairwise gram refine diffab api ingest; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2,-Wall \
      --cudart static "$@" -c $f.cu -o _build_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o libfununifrac_$name.so _build_$name/*.o --cudart static -gencode arch=compute_100a,code=sm_100a
echo "built fununifrac_b200/csrc/libfununifrac_$name.so"
