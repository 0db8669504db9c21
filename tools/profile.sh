#!/bin/bash
# ncu evidence for one round (run under gpurun): launch list of a few bench steps + full captures of the top kernels.
#   bash tools/profile.sh r1
# The bench is run with eager launches (--no-cuda-graphs): same kernels, and every launch is visible to ncu one by one.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --global-batch 128 --micro-batch 128 --no-cpu-baseline --no-eval --no-cuda-graphs"
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_$TAG.csv $BENCH > $OUT/launches_$TAG.stdout 2>&1
for K in ${KERNELS:-attn_tc_bwd_kernel attn_fwd_small_kernel gemm_tn_kernel wgrad_kernel qk_norm_rope_bwd_kernel rmsnorm_bwd_kernel}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 2 -f -o $OUT/prof_${TAG}_$K $BENCH > $OUT/prof_${TAG}_$K.stdout 2>&1
done
# the K1 kernels at the long-history size (tools/embed_bench.py)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'embed_route_kernel|emb_reduce_kernel' -s 2 -c 2 -f -o $OUT/prof_${TAG}_embed python tools/embed_bench.py --iters 3 > $OUT/prof_${TAG}_embed.stdout 2>&1
ls -la $OUT | tail -20
