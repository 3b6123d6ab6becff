#!/bin/bash
# build variants of rqp_single.cu (compile-time switches) into build/variants/librqp_<tag>.so
cd "$(dirname "$0")/../reluqp-py_b200"
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v"
build() { tag=$1; shift
  nvcc $FLAGS "$@" -c csrc/rqp_single.cu -o build/variants/single_$tag.o 2> build/variants/single_$tag.log &&
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/librqp_$tag.so build/variants/single_$tag.o build/rqp_batched.o build/rqp_batched_tc.o build/rqp_abi.o -lcudart_static -lpthread -ldl -lrt
  grep -A2 "IdLi2ELi256ELb1" build/variants/single_$tag.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo " <- $tag"; }
for v in "$@"; do
  case $v in
    s0a0) build s0a0 -DRQP_V_STAGE=0 -DRQP_V_AG=0 & ;;
    s1a0) build s1a0 -DRQP_V_STAGE=1 -DRQP_V_AG=0 & ;;
    s0a1) build s0a1 -DRQP_V_STAGE=0 -DRQP_V_AG=1 & ;;
    s1a1) build s1a1 -DRQP_V_STAGE=1 -DRQP_V_AG=1 & ;;
    t00) build t00 -DRQP_V_CHKTIME=1 & ;;
    t11) build t11 -DRQP_V_CHKTIME=1 -DRQP_V_STAGE=1 -DRQP_V_AG=1 & ;;
    s2a1) build s2a1 -DRQP_V_STAGE=2 -DRQP_V_AG=1 & ;;
  esac
done
wait
