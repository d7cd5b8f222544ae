#!/bin/bash
# Builds the three stand-alone microbenchmarks next to their sources (binaries are git-ignored; they travel to the GPU box).
#   pipe_probe   issue rate of the ALU / FMA-pipe instructions the matcher's drain is made of, and of candidate replacements
#   tmem_probe   tcgen05.ld .pack::16b / tcgen05.st .unpack::16b semantics, exact accumulation onto a magic constant, LDTM / STTM rates
#   drain_probe  the float drain's 16-column slice in isolation (what the ALU pipe sustains for exactly that instruction mix)
set -e
cd "$(dirname "$0")"
for p in pipe_probe tmem_probe drain_probe; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I ../../vslam_b200/csrc -o $p $p.cu -lcuda
done
