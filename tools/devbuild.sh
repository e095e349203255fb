#!/bin/bash
# Kernel-tuning build: libvmk_dev.so with only the 1024 and 8192 size families (seconds instead of minutes of ptxas).
# Use with VMK_LIB=$PWD/cfd_julia_b200/libvmk_dev.so (cfd_julia_b200/_lib.py).  Extra nvcc flags: "$@".
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -Xcompiler -fPIC -shared \
     -DVMK_DEV_SIZES "$@" -o cfd_julia_b200/libvmk_dev.so cfd_julia_b200/csrc/vmk.cu
