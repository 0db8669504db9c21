// tcgen05.mma throughput by operand flavour (one CTA per SM, one issuing thread, operands resident in shared memory).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I gamer_b200/csrc -o tools/microbench/umma tools/microbench/umma.cu
// Prints cycles per MMA instruction for: K-major x K-major N=128 (QK^T), MN x MN N=64 (P^T dO), K x MN N=64 (dS K),
// TMEM-A x MN N=64 (P V), with and without concurrent shared-memory store traffic from 8 other warps.
#include <cstdio>
#include "sm100_ptx.cuh"
using namespace sm100;

__device__ __forceinline__ uint64_t dk(uint32_t a) { return umma_desc_sw128(a, 16, 1024); }
__device__ __forceinline__ uint64_t dmn(uint32_t a, uint32_t lbo) { return umma_desc_sw128(a, lbo, 1024); }

template <int MODE, bool NOISE, bool TLD = false, int CEVERY = 0>
__global__ void __launch_bounds__(320, 1) bench(long long* out, int iters) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint64_t dummy[4];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&dummy[i], 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 65536);
    if (warp == 0) {
        long long t0 = clock64();
        uint32_t par = 0;
        for (int it = 0; it < iters; ++it) {
            if (elect_one()) {
                if (MODE == 0) {          // S = Q K^T: 128x128x64, both K-major
                    constexpr uint32_t id = umma_idesc_bf16(128, 128, 0, 0);
                    const uint64_t a = dk(sa), b = dk(sb);
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16(tm + (r & 1) * 128, a + 2 * k + r * 1024, b + 2 * k, id, k != 0);
                } else if (MODE == 1) {   // dV += P^T dO: 128x64x128, both MN-major
                    constexpr uint32_t id = umma_idesc_bf16(128, 64, 1, 1);
                    const uint64_t a = dmn(sa, 16384), b = dmn(sb, 16384);
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            umma_bf16(tm + 256 + r * 64, a + k * 128, b + k * 128, id, 1);
                            if (CEVERY && ((r * 8 + k + 1) % CEVERY) == 0) umma_commit(&dummy[(r * 8 + k) / CEVERY % 4]);
                        }
                } else if (MODE == 2) {   // dQ = dS K: 128x64x128, A K-major (two halves), B MN-major
                    constexpr uint32_t id = umma_idesc_bf16(128, 64, 0, 1);
                    const uint64_t a = dk(sa), b = dmn(sb, 16384);
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int k = 0; k < 8; ++k) umma_bf16(tm + 384 + r * 64, a + (k >> 2) * 1024 + (k & 3) * 2, b + k * 128, id, 1);
                } else {                  // O += P V: A in TMEM, B MN-major, 128x64x64 (x4 to make 16 MMAs)
                    constexpr uint32_t id = umma_idesc_bf16(128, 64, 0, 1);
                    const uint64_t b = dmn(sb, 8192);
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_bf16_ts(tm + 128 + (r & 1) * 64, tm + k * 8, b + k * 128, id, 1);
                }
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, par);
            par ^= 1;
        }
        long long t1 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    } else if (NOISE && TLD && warp >= 2) {
        // 8 warps streaming tcgen05.ld of fp32 accumulator columns (the softmax warps' S / dP reads)
        const uint32_t t = tm + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) & 1) * 64;
        uint32_t acc = 0;
        for (int it = 0; it < iters * 4; ++it) {
            uint32_t r[32];
            tmem_ld_32x32(t + (it & 1) * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 32; ++q) acc ^= r[q];
        }
        if (acc == 0x12345678u) out[1] = acc;
    } else if (NOISE && warp >= 2) {
        // 8 warps streaming 16-byte stores into a third region (the softmax warps' P / dS writes)
        const uint32_t dst = smem_u32(smem + 131072) + (threadIdx.x - 64) * 128;
        for (int it = 0; it < iters * 8; ++it) {
#pragma unroll
            for (int q = 0; q < 8; ++q) sts128(dst + ((q ^ (threadIdx.x & 7)) << 4), it, q, it, q);
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

template <int MODE, bool NOISE, bool TLD = false, int CEVERY = 0>
void run(const char* name, long long* d_out) {
    const int smem = 200 * 1024, iters = 200;
    cudaFuncSetAttribute(bench<MODE, NOISE, TLD, CEVERY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    bench<MODE, NOISE, TLD, CEVERY><<<148, 320, smem>>>(d_out, iters);
    bench<MODE, NOISE, TLD, CEVERY><<<148, 320, smem>>>(d_out, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%-44s noise=%d tmem_ld=%d  %7.1f cycles per MMA (16 MMAs + commit + wait per iteration)  [%s]\n", name, (int)NOISE, (int)TLD,
           (double)h / (iters * 16.0), cudaGetErrorString(e));
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 16);
    run<0, false>("K-major x K-major 128x128x16 (S, dP)", d_out);
    run<0, true>("K-major x K-major 128x128x16 (S, dP)", d_out);
    run<1, false>("MN x MN 128x64x16 (dV, dK)", d_out);
    run<1, true>("MN x MN 128x64x16 (dV, dK)", d_out);
    run<2, false>("K x MN 128x64x16 (dQ)", d_out);
    run<2, true>("K x MN 128x64x16 (dQ)", d_out);
    run<0, true, true>("K-major x K-major 128x128x16 (S, dP)", d_out);
    run<1, true, true>("MN x MN 128x64x16 (dV, dK)", d_out);
    run<2, true, true>("K x MN 128x64x16 (dQ)", d_out);
    run<3, true, true>("TMEM x MN 128x64x16 (PV)", d_out);
    run<1, false, false, 8>("MN x MN 128x64x16, commit every 8", d_out);
    run<1, false, false, 4>("MN x MN 128x64x16, commit every 4", d_out);
    run<1, false, false, 2>("MN x MN 128x64x16, commit every 2", d_out);
    run<3, false>("TMEM x MN 128x64x16 (PV)", d_out);
    run<3, true>("TMEM x MN 128x64x16 (PV)", d_out);
    return 0;
}
