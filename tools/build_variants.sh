#!/bin/bash
# Tuning variants of the library (loads in flight per thread, min CTAs/SM) -> smart_b200/lib/variants/
set -e
cd "$(dirname "$0")/../smart_b200/csrc"
mkdir -p ../lib/variants ../../build/var
gcc -O2 -fPIC -std=c11 -D_POSIX_C_SOURCE=200809L -I../../include -I. -c -o ../../build/var/prep.o prep.c
for v in "2 1" "2 5" "2 6" "3 1" "4 1" "4 4" "1 8"; do
  set -- $v
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include -I. \
       -DSB_U=$1 -DSB_MINB=$2 -c -o ../../build/var/api_$1_$2.o api.cu
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/variants/libsmartb200_u$1_b$2.so ../../build/var/api_$1_$2.o ../../build/var/prep.o
  echo built u$1 b$2
done
