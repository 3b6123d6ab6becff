#!/bin/bash
# Build variants of rqp_single.cu (compile-time switches) into build/variants/librqp_<tag>.so; tools/variants_run.sh
# then runs the bench once per variant on the GPU box.  Current switch: RQP_V_CHKTIME=1 (the phase counters
# time the steps of the residual check: staging, row products, CTA reduction + publish, all-gather, logic).
cd "$(dirname "$0")/../reluqp-py_b200"
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v"
build() { tag=$1; shift
  nvcc $FLAGS "$@" -c csrc/rqp_single.cu -o build/variants/single_$tag.o 2> build/variants/single_$tag.log &&
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/librqp_$tag.so build/variants/single_$tag.o build/rqp_batched.o build/rqp_batched_tc.o build/rqp_abi.o -lcudart_static -lpthread -ldl -lrt
  grep -A2 "IdLi2ELi256ELb1" build/variants/single_$tag.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo " <- $tag"; }
build base &
build halfpoll -DRQP_V_HALFPOLL=1 &
wait
