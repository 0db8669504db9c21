// Microbenchmark for the sparse embedding gradient (emb_reduce_kernel's design space):
//   A: gather rows in id-sorted (= random) order, R rows in flight per warp, sum in registers  -> random 512 B DRAM reads
//   B: stream rows in storage order (coalesced) and red.global.add.v4.f32 them into the L2-resident table
//   C: stream rows, one red per (row, lane) after a warp-level pre-sum of EQUAL ids (match_any)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o embgrad embgrad.cu ; ./embgrad [rows] [vocab]
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <vector>

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e = (x);                                                        \
        if (e != cudaSuccess) {                                                     \
            printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));        \
            exit(1);                                                                \
        }                                                                           \
    } while (0)

constexpr int H = 256;  // bf16 columns per row: 512 B

__device__ __forceinline__ void acc8(float* a, uint4 r) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 f = __bfloat1622float2(p[k]);
        a[2 * k] += f.x;
        a[2 * k + 1] += f.y;
    }
}

// A: one warp per chunk of 64 sorted rows; R rows in flight
template <int R>
__global__ void gather_sum(const __nv_bfloat16* __restrict__ dx, const int* __restrict__ order, int n_chunks,
                           float* __restrict__ out, const int* __restrict__ ids = nullptr, int vec = 0) {
    const int lane = threadIdx.x & 31;
    for (int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); ch < n_chunks; ch += gridDim.x * (blockDim.x >> 5)) {
        const int i0 = order[ch * 64 + lane], i1 = order[ch * 64 + 32 + lane];
        float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
        for (int i = 0; i < 64; i += R) {
            uint4 r[R];
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const int k = i + u;
                const int row = __shfl_sync(0xffffffffu, k < 32 ? i0 : i1, k & 31);
                r[u] = __ldcs(reinterpret_cast<const uint4*>(dx + (long long)row * H) + lane);
            }
#pragma unroll
            for (int u = 0; u < R; ++u) acc8(a, r[u]);
        }
        const int trow = ids ? ids[__shfl_sync(0xffffffffu, i0, 0)] : (ch & 1023);
        float* o = out + (long long)trow * H + lane * 8;
        if (vec) {
            atomicAdd(reinterpret_cast<float4*>(o), make_float4(a[0], a[1], a[2], a[3]));
            atomicAdd(reinterpret_cast<float4*>(o) + 1, make_float4(a[4], a[5], a[6], a[7]));
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(o + k, a[k]);
        }
    }
}

// B: stream, vector reds
template <int R>
__global__ void stream_red(const __nv_bfloat16* __restrict__ dx, const int* __restrict__ ids, long long rows,
                           float* __restrict__ table) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long r0 = w * 32; r0 < rows; r0 += nw * 32) {
        const int id = (r0 + lane < rows) ? ids[r0 + lane] : -1;
#pragma unroll 1
        for (int i = 0; i < 32; i += R) {
            uint4 r[R];
#pragma unroll
            for (int u = 0; u < R; ++u)
                r[u] = (r0 + i + u < rows) ? __ldcs(reinterpret_cast<const uint4*>(dx + (r0 + i + u) * H) + lane) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < R; ++u) {
                const int v = __shfl_sync(0xffffffffu, id, i + u);
                if (v < 0) continue;
                float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                acc8(a, r[u]);
                float4* o = reinterpret_cast<float4*>(table + (long long)v * H + lane * 8);
                atomicAdd(o, make_float4(a[0], a[1], a[2], a[3]));
                atomicAdd(o + 1, make_float4(a[4], a[5], a[6], a[7]));
            }
        }
    }
}

// D: column-slice privatisation: CTA (slice s of 8, partition p) keeps table[:, 32 s .. 32 s + 31] in shared memory
// (vocab x 32 fp32), streams the 64 B slice of every row of its partition, shared CAS-adds, one flush with reds.
__global__ void slice_smem(const __nv_bfloat16* __restrict__ dx, const int* __restrict__ ids, long long rows, int vocab,
                           float* __restrict__ table) {
    extern __shared__ float sm[];
    const int slice = blockIdx.x & 7, part = blockIdx.x >> 3, parts = gridDim.x >> 3;
    for (int i = threadIdx.x; i < vocab * 32; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const long long per = (rows + parts - 1) / parts;
    const long long rb = part * per, re = min(rows, rb + per);
    const int sub = threadIdx.x & 3;  // 16 B piece of the 64 B slice
    for (long long r = rb + (threadIdx.x >> 2); r < re; r += (blockDim.x >> 2) * 4) {
        uint4 v[4];
        int id[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long rr = r + (long long)u * (blockDim.x >> 2);
            id[u] = rr < re ? ids[rr] : -1;
            v[u] = rr < re ? __ldcs(reinterpret_cast<const uint4*>(dx + rr * H + slice * 32) + sub) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (id[u] < 0) continue;
            float a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            acc8(a, v[u]);
            float* o = sm + id[u] * 32 + sub * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(o + ((k + sub * 2) & 7), a[(k + sub * 2) & 7]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < vocab * 32; i += blockDim.x) {
        const float x = sm[i];
        if (x != 0.f) atomicAdd(table + (long long)(i >> 5) * H + slice * 32 + (i & 31), x);
    }
}

template <class F>
static float time_ms(F f, int iters = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    f();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / iters;
}

int main(int argc, char** argv) {
    const long long rows = argc > 1 ? atoll(argv[1]) : 512LL * 2505;
    const int vocab = argc > 2 ? atoi(argv[2]) : 1041;
    printf("rows %lld vocab %d (%.1f MB of bf16 rows)\n", rows, vocab, rows * H * 2 / 1e6);
    std::vector<int> ids(rows), order(rows);
    srand(1);
    const int skew = argc > 3 ? atoi(argv[3]) : 0;     // 1: every 5th row carries one of 3 behaviour ids (the SMB layout)
    for (long long i = 0; i < rows; ++i) ids[i] = (skew && i % 5 == 0) ? rand() % 3 : 3 + rand() % (vocab - 3);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ids[a] < ids[b]; });
    if (argc > 4 && atoi(argv[4])) {    // order WITHIN an id shuffled (what an atomics-based counting sort produces at worst)
        long long b = 0;
        while (b < rows) {
            long long e = b;
            while (e < rows && ids[order[e]] == ids[order[b]]) ++e;
            for (long long i = e - 1; i > b; --i) std::swap(order[i], order[b + rand() % (i - b + 1)]);
            b = e;
        }
        printf("order within an id: shuffled\n");
    }
    __nv_bfloat16* dx;
    int *d_ids, *d_order;
    float* table;
    CK(cudaMalloc(&dx, rows * H * 2));
    CK(cudaMemset(dx, 0x3c, rows * H * 2));
    CK(cudaMalloc(&d_ids, rows * 4));
    CK(cudaMalloc(&d_order, (rows + 64) * 4));
    CK(cudaMemset(d_order, 0, (rows + 64) * 4));
    CK(cudaMalloc(&table, (size_t)std::max(vocab, 1024) * H * 4));
    CK(cudaMemcpy(d_ids, ids.data(), rows * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_order, order.data(), rows * 4, cudaMemcpyHostToDevice));
    const double bytes = rows * (H * 2.0 + 4.0);
    const int n_chunks = (int)(rows / 64);
    auto report = [&](const char* name, float ms) { printf("%-34s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, bytes / ms / 1e6); };
    for (int blocks : {148 * 3, 148 * 6, 148 * 8}) {
        char nm[64];
        snprintf(nm, 64, "A gather R=4  grid %d x256", blocks);
        report(nm, time_ms([&] { gather_sum<4><<<blocks, 256>>>(dx, d_order, n_chunks, table); }));
        snprintf(nm, 64, "A gather R=8  grid %d x256", blocks);
        report(nm, time_ms([&] { gather_sum<8><<<blocks, 256>>>(dx, d_order, n_chunks, table); }));
        snprintf(nm, 64, "A gather R=16 grid %d x256", blocks);
        report(nm, time_ms([&] { gather_sum<16><<<blocks, 256>>>(dx, d_order, n_chunks, table); }));
    }
    report("A R=8 atomics on the chunk's id row", time_ms([&] { gather_sum<8><<<148 * 8, 256>>>(dx, d_order, n_chunks, table, d_ids, 0); }));
    report("A R=8 same, red.v4", time_ms([&] { gather_sum<8><<<148 * 8, 256>>>(dx, d_order, n_chunks, table, d_ids, 1); }));
    // sequential order through the same gather kernel: the DRAM-friendly bound of design A
    std::iota(order.begin(), order.end(), 0);
    CK(cudaMemcpy(d_order, order.data(), rows * 4, cudaMemcpyHostToDevice));
    report("A gather R=8 SEQUENTIAL order", time_ms([&] { gather_sum<8><<<148 * 8, 256>>>(dx, d_order, n_chunks, table); }));
    for (int blocks : {148 * 4, 148 * 8}) {
        char nm[64];
        snprintf(nm, 64, "B stream+red.v4 R=4 grid %d", blocks);
        report(nm, time_ms([&] { stream_red<4><<<blocks, 256>>>(dx, d_ids, rows, table); }));
        snprintf(nm, 64, "B stream+red.v4 R=8 grid %d", blocks);
        report(nm, time_ms([&] { stream_red<8><<<blocks, 256>>>(dx, d_ids, rows, table); }));
    }
    const int smem = vocab * 32 * 4;
    CK(cudaFuncSetAttribute(slice_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int parts : {18, 37}) {
        char nm[64];
        snprintf(nm, 64, "D slice smem parts %d x1024", parts);
        report(nm, time_ms([&] { slice_smem<<<parts * 8, 1024, smem>>>(dx, d_ids, rows, vocab, table); }));
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
