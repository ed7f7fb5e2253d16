#!/bin/bash
B=tools/mma_bench/mma_bench
R=2000
echo "== CTAs per SM (tf32 none M=128 N=16)"
for C in 1 2 4; do $B 0 128 16 0 0 0 2080 128 256 128 0 1 $R 148 1 $C 1; done
echo "== issuing warps per CTA (own accumulators)"
for W in 1 2 4; do $B 0 128 16 0 0 0 2080 128 256 128 0 1 $R 148 1 1 $W; done
echo "== CTAs per SM, N=64 and N=256(bf16 sw128)"
for C in 1 2 4; do $B 0 128 64 0 0 0 2080 128 1024 128 0 1 $R 148 1 $C 1; done
for C in 1 2; do $B 1 128 256 0 0 2 16 1024 16 1024 0 1 $R 148 1 $C 1; done
echo "== 4 CTAs/SM x 2 warps"
$B 0 128 16 0 0 0 2080 128 256 128 0 1 $R 148 1 4 2
