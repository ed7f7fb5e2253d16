#!/bin/bash
# kind M N amaj bmaj layout lboA sboA lboB sboB ts nacc reps
B=tools/mma_bench/mma_bench
R=2000
echo "== tf32 K-major no-swizzle (forward today), N sweep"
for N in 16 32 64 128 256; do $B 0 128 $N 0 0 0 2080 128 $((N*16)) 128 0 1 $R; done
echo "== tf32 K-major no-swizzle M=64"
for N in 16 64; do $B 0 64 $N 0 0 0 2080 128 $((N*16)) 128 0 1 $R; done
echo "== tf32 K-major SW128 (sbo 1024)"
for N in 16 32 64 128 256; do $B 0 128 $N 0 0 2 16 1024 16 1024 0 1 $R; done
echo "== tf32 K-major SW64 (sbo 512), SW32 (sbo 256)"
for N in 16 64; do $B 0 128 $N 0 0 4 16 512 16 512 0 1 $R; $B 0 128 $N 0 0 6 16 256 16 256 0 1 $R; done
echo "== tf32 TS mode (A in TMEM), B no-swizzle / SW128"
for N in 16 32 64 128; do $B 0 128 $N 0 0 0 2080 128 $((N*16)) 128 1 1 $R; $B 0 128 $N 0 0 2 16 1024 16 1024 1 1 $R; done
echo "== bf16 MN-major no-swizzle (wgrad today)"
for M in 64 128; do for N in 16 32 64; do $B 1 $M $N 1 1 0 256 128 256 128 0 1 $R; done; done
echo "== bf16 MN-major SW128 / SW64 / SW32 (lbo = MN group stride, sbo = 8-row K group stride)"
for N in 16 64; do $B 1 64 $N 1 1 2 8192 1024 8192 1024 0 1 $R; $B 1 64 $N 1 1 4 4096 512 4096 512 0 1 $R; $B 1 64 $N 1 1 6 2048 256 2048 256 0 1 $R; done
echo "== bf16 K-major no-swizzle / SW128, M=64/128"
for M in 64 128; do for N in 16 64; do $B 1 $M $N 0 0 0 2080 128 $((N*16)) 128 0 1 $R; $B 1 $M $N 0 0 2 16 1024 16 1024 0 1 $R; done; done
echo "== bf16 TS mode"
for N in 16 64; do $B 1 128 $N 0 0 0 2080 128 $((N*16)) 128 1 1 $R; $B 1 128 $N 0 1 0 256 128 256 128 1 1 $R; done
echo "== independent accumulators (tf32 no-swizzle N=16)"
for A in 1 2 4 8; do $B 0 128 16 0 0 0 2080 128 256 128 0 $A $R; done
echo "== one CTA only (no cross-SM effects)"
$B 0 128 16 0 0 0 2080 128 256 128 0 1 $R 1
$B 0 128 16 0 0 2 16 1024 16 1024 0 1 $R 1
