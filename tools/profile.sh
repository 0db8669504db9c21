#!/bin/bash
# ncu evidence for one round (run under gpurun): launch list of one bench step + full captures of the top kernels.
#   bash tools/profile.sh r2
# The bench is run with eager launches (--no-cuda-graphs): same kernels, and every launch is visible to ncu one by one.
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -f"
# every launch of one headline step (global batch 1024 = one pass of 1024 rows) with its device time (cold-cache,
# serialised: compare SHARES, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-eval --no-cuda-graphs > $OUT/launches_$TAG.stdout 2>&1
# attention kernels at the bench's micro-batch (1024 rows, L = 505, dropout 0.2): self (kind 0) and cross (kind 1)
for K in 0 1; do
  timeout 300 $NCU -k regex:'attn_bwd_kernel|attn_fwd_kernel' -s 2 -c 2 -o $OUT/prof_${TAG}_attn_k$K \
    python tools/attn_one.py --kind $K --p 0.2 --batch 1024 --iters 2 --bench-levels > $OUT/prof_${TAG}_attn_k$K.stdout 2>&1
done
# decode kernels at the eval shape (256 users x 20 beams, 501-token prompts)
timeout 300 $NCU -k regex:'attn_decode_kernel|beam_step_kernel' -s 26 -c 4 -o $OUT/prof_${TAG}_decode \
  python tools/eval_one.py --users 256 --iters 2 > $OUT/prof_${TAG}_decode.stdout 2>&1
# the K1 kernels at the long-history size (tools/embed_bench.py)
timeout 300 $NCU -k regex:'embed_route_kernel|emb_reduce_kernel' -s 2 -c 2 -o $OUT/prof_${TAG}_embed \
  python tools/embed_bench.py --iters 3 > $OUT/prof_${TAG}_embed.stdout 2>&1
ls -la $OUT | tail -12
# DRAM traffic of every GEMM launch of one eager headline step (bench.py's roofline.traffic for the GEMM entry points)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:'gemm_tn_kernel|wgrad_kernel' -s 168 -c 168 --csv --log-file $OUT/gemm_traffic_$TAG.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-eval --no-cuda-graphs > $OUT/gemm_traffic_$TAG.stdout 2>&1
# the memory-bound backward kernels inside the step
timeout 600 $NCU -k regex:'qk_norm_rope_bwd_kernel|rmsnorm_bwd_kernel|swiglu_bwd_kernel|qk_norm_rope_fwd_kernel' -s 30 -c 10 \
  -o $OUT/prof_${TAG}_elem python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-baseline --no-eval --no-cuda-graphs \
  > $OUT/prof_${TAG}_elem.stdout 2>&1
ls -la $OUT | tail -12

