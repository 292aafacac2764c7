#!/bin/bash
# Development aid: build the CUDA library of another git revision next to the current one, for A/B timing on one box:
#   scripts/build_rev.sh <rev>   ->  content-aware-gan-compression_b200/lib/ab_<rev>.so   (load it with CAGC_LIB=<path>)
set -e
rev=$1
root=$(cd "$(dirname "$0")/.." && pwd)
pkg=content-aware-gan-compression_b200
tmp=$(mktemp -d)
git -C "$root" archive "$rev" $pkg/csrc include | tar -x -C "$tmp"
srcs=$(cd "$tmp/$pkg/csrc" && ls *.cu | sed "s|^|$tmp/$pkg/csrc/|")
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
     -I "$tmp/include" -I "$tmp/$pkg/csrc" $srcs -o "$root/$pkg/lib/ab_$rev.so"
rm -rf "$tmp"
echo "$root/$pkg/lib/ab_$rev.so"
