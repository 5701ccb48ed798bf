#!/bin/bash
# Build an experimental copy of the library with extra nvcc flags (e.g. -DMG_REC=1):
#   scripts/build_variant.sh NAME [flags...]  ->  remora_b200/lib/variants/librb200_NAME.so
# Load it with RB200_LIB=<that path> (remora_b200/_native.py).  Only rb200_mega.cu is recompiled; the other
# objects come from the regular build (python -m remora_b200.build_native).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p remora_b200/lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -O2 \
  -Wno-deprecated-gpu-targets -I include -I remora_b200/csrc "$@" -c ${MEGA_SRC:-remora_b200/csrc/rb200_mega.cu} \
  -o remora_b200/lib/variants/mega_$name.o
objs=$(ls remora_b200/lib/*.o | grep -v rb200_mega.o)
nvcc -shared -o remora_b200/lib/variants/librb200_$name.so $objs remora_b200/lib/variants/mega_$name.o -lcudart
echo remora_b200/lib/variants/librb200_$name.so
