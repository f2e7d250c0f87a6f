#!/bin/sh
# Instrumented build (-DDOGSTEP_TRACE) -> scripts/microbench/libdogstep_trace.so, used through DOGSTEP_LIB by scripts/prof_*.py:
#   det-MADN play kernel: per-round timeline of one CTA (DOGSTEP_PLAY_TRACE=<cta>; DOGSTEP_PENTER=<n>, DOGSTEP_ROUND=<n> override
#   the draw-ahead threshold and the round length), DOG play kernel: timeline of CTA 0 + phase cycles of a lone game
#   (DOGSTEP_DOG_TRACE=1), TicTacToe search: rollout and phase counters (DOGSTEP_TTT_TRACE=1)
set -e
cd "$(dirname "$0")/../exploring-muzero-on-dog_b200/csrc"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --fmad=false -DDOGSTEP_TRACE"
for f in madn_kernels mcts_kernels dog_kernels; do nvcc $FLAGS -c $f.cu -o /tmp/${f}_trace.o; done
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../../scripts/microbench/libdogstep_trace.so /tmp/madn_kernels_trace.o \
  /tmp/mcts_kernels_trace.o /tmp/dog_kernels_trace.o abi_common.o replay_kernels.o ttt_kernels.o -lcudart
