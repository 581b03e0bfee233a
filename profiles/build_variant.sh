#!/bin/bash
# dev helper: profiles/build_variant.sh <name> <file.cu> "<-D flags>"  ->  variants/libsrj_<name>.so
# (all other objects come from the regular build; select at run time with SRJ_B200_LIB=<path>)
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; flags=$3
mkdir -p variants
P=spark-rapids-jni_b200
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -ccbin /usr/bin/g++ --expt-relaxed-constexpr $flags -c $P/csrc/$src -o variants/$name.o
objs=""
for o in $P/build/*.o; do [ "$(basename $o)" = "${src%.cu}.o" ] || objs="$objs $o"; done
nvcc --shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o variants/libsrj_$name.so $objs variants/$name.o
echo variants/libsrj_$name.so
