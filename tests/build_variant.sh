#!/bin/bash
# Build a variant of libflairb200.so with extra nvcc defines (A/B runs inside one GPU job, loaded with FB_LIB_PATH):
#   tests/build_variant.sh flair-1_b200/libflairb200_alt.so -DFB_ACC_DEEP=2
set -e
out=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
for f in conv_igemm conv_halo elementwise api comm host_codec; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -I "$root/include" \
       -c "$root/flair-1_b200/csrc/$f.cu" -o "$tmp/$f.o" 2> >(grep -v "warning\|Remark\|^$\|declared but never\|\^" >&2) &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a --shared -Xcompiler -fPIC -cudart static -ldl -o "$out" "$tmp"/*.o
rm -rf "$tmp"
echo "$out"
