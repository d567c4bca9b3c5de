#!/bin/bash
# builds ablation variants of node_tc.cu (PILE_ABL bit mask) into abl/ -- measurement aid
set -e
cd "$(dirname "$0")/../dyn_res_pile_manip_b200/csrc"
make -s
for a in "$@"; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DPILE_ABL=$a -c node_tc.cu -o ../../abl/node_tc_$a.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../abl/libpilegnn_abl$a.so api.o nbr.o fwd.o reward.o bwd.o edge_tc.o ../../abl/node_tc_$a.o edge_tmem.o bwd_tc.o -cudart static
done
