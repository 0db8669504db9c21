// tcgen05.ld throughput: NW warps of one CTA per SM stream 32x32b.x32 loads (4 KB per warp instruction).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I gamer_b200/csrc -o tools/microbench/tmemld tools/microbench/tmemld.cu
#include <cstdio>
#include "sm100_ptx.cuh"
using namespace sm100;

template <int NW, int DEPTH>
__global__ void __launch_bounds__(NW * 32, 1) bench(long long* out, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t t = tm + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t r[DEPTH][32];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) tmem_ld_32x32(t + (d & 1) * 32, r[d]);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
#pragma unroll
            for (int q = 0; q < 32; ++q) acc ^= r[d][q];
    }
    __syncthreads();
    long long t1 = clock64();
    if (acc == 0x12345678u) out[1] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

template <int NW, int DEPTH>
void run(long long* d_out) {
    const int iters = 2000;
    bench<NW, DEPTH><<<148, NW * 32>>>(d_out, iters);
    bench<NW, DEPTH><<<148, NW * 32>>>(d_out, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)iters * DEPTH * NW * 4096.0;
    printf("warps=%2d loads in flight=%d: %6.1f B/clk/SM, %6.1f cycles per warp-load  [%s]\n", NW, DEPTH, bytes / h,
           (double)h / (iters * DEPTH), cudaGetErrorString(e));
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 16);
    run<4, 1>(d_out); run<4, 2>(d_out); run<8, 1>(d_out); run<8, 2>(d_out); run<16, 1>(d_out); run<16, 2>(d_out);
    return 0;
}
