#!/bin/bash
# build_variant.sh <name> <extra nvcc defines...>  -> build/variants/<name>.so  (tuning sweeps; the product build is __graft_entry__.build())
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true --extended-lambda \
  -Xcompiler -fPIC -shared -Xptxas -v -ccbin g++ "$@" exadem_b200/csrc/edb_lib.cu -o build/variants/$name.so -ldl 2> build/variants/$name.log
grep -A3 "k_contact_sphere_ppILi0\|k_gatherILi2" build/variants/$name.log | grep "Used\|spill" 
