#!/bin/bash
# builds ablation variants of edge_tmem.cu (PILE_ABL bit mask) into abl/ -- measurement aid
set -e
cd "$(dirname "$0")/../dyn_res_pile_manip_b200/csrc"
for a in "$@"; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DPILE_ABL=$a -c edge_tmem.cu -o ../../abl/edge_tmem_$a.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../abl/libpilegnn_eabl$a.so api.o nbr.o fwd.o reward.o bwd.o edge_tc.o node_tc.o ../../abl/edge_tmem_$a.o bwd_tc.o -cudart static
done
