#!/bin/bash
# development: build a variant of libnnb.so that differs from the product library only in the d = 30 tensor-core unit
#   scripts/dev/build_variant.sh lib_v1 "-DNNB_TC_WAIT_HINT"     ->  nnest_b200/lib_v1/libnnb.so
# (needs an up-to-date product build: the other objects are taken from nnest_b200/lib/obj)
set -e
cd "$(dirname "$0")/../.."
dir=nnest_b200/$1; shift
mkdir -p $dir
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v $@ \
  -c nnest_b200/csrc/nnb_tc_d30.cu -o $dir/nnb_tc_d30.o 2> $dir/ptxas_d30.log
objs=$(ls nnest_b200/lib/obj/*.o | grep -v nnb_tc_d30.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $dir/libnnb.so $objs $dir/nnb_tc_d30.o
grep -A1 "mcmc_tc_kernelILi0ELi1ELi30" $dir/ptxas_d30.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo " -> $dir/libnnb.so"
