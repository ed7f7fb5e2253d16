#!/bin/bash
B=tools/mma_bench/mma_bench
R=2000
echo "== elect_one pattern: tf32 K-major no-swizzle N sweep"
for N in 16 32 64 128 256; do $B 0 128 $N 0 0 0 2080 128 $((N*16)) 128 0 1 $R 148 1; done
echo "== elect: tf32 SW128"
for N in 16 64 256; do $B 0 128 $N 0 0 2 16 1024 16 1024 0 1 $R 148 1; done
echo "== elect: tf32 TS"
for N in 16 64 256; do $B 0 128 $N 0 0 0 2080 128 $((N*16)) 128 1 1 $R 148 1; done
echo "== elect: M=64"
for N in 16 64; do $B 0 64 $N 0 0 0 2080 128 $((N*16)) 128 0 1 $R 148 1; done
echo "== elect: bf16 MN-major none M=64/128"
for M in 64 128; do for N in 16 64; do $B 1 $M $N 1 1 0 256 128 256 128 0 1 $R 148 1; done; done
echo "== elect: bf16 K-major SW128"
for N in 16 64 256; do $B 1 128 $N 0 0 2 16 1024 16 1024 0 1 $R 148 1; done
echo "== elect: nacc"
for A in 1 2 4; do $B 0 128 16 0 0 0 2080 128 256 128 0 $A $R 148 1; done
