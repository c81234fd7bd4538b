#!/bin/bash
# usage: tools/build_variant.sh name [-DMACRO=VALUE ...]   -> build/ab/lib_<name>.so  (kernel-tuning experiments)
name=$1; shift
mkdir -p build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -Xptxas=-v -ccbin /usr/bin/g++ \
  "$@" -o build/ab/lib_$name.so realtime-collision-detection_b200/csrc/rcd_api.cu > build/ab/$name.log 2>&1 || { echo "$name FAILED"; tail -5 build/ab/$name.log; exit 1; }
echo "$name: $(grep -A3 'k_onesweep_passILb0' build/ab/$name.log | grep Used)"
