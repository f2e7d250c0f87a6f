#!/bin/sh
# instrumented build of the det-MADN play kernel (per-round timeline of one CTA; DOGSTEP_PLAY_TRACE=<cta>, DOGSTEP_PENTER=<n>)
set -e
cd "$(dirname "$0")/../exploring-muzero-on-dog_b200/csrc"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --fmad=false -DDOGSTEP_TRACE -c madn_kernels.cu -o /tmp/madn_trace.o
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --fmad=false -DDOGSTEP_TRACE -c mcts_kernels.cu -o /tmp/mcts_trace.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../scripts/microbench/libdogstep_trace.so /tmp/madn_trace.o abi_common.o dog_kernels.o /tmp/mcts_trace.o replay_kernels.o ttt_kernels.o -lcudart
