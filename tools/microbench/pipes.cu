// Instruction-throughput microbenchmark (B200): cycles per warp-instruction per SMSP for the pipes the softmax warps use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
template <int MODE>
__global__ void k(float* out, long long* cyc, float a, float b) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = a + threadIdx.x * 0.001f + i;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) x[i] = fmaf(x[i], a, b);                                       // FFMA
            if (MODE == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));        // MUFU.EX2
            if (MODE == 2) x[i] = fmaxf(x[i], b + i);                                     // FMNMX
            if (MODE == 3) x[i] = (x[i] < b + i) ? x[i] : a;                              // FSETP + FSEL
            if (MODE == 4) {                                                              // MUFU + 2 FFMA
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
                x[i] = fmaf(x[i], a, b);
                x[i] = fmaf(x[i], a, b);
            }
            if (MODE == 5) {                                                              // MUFU + FSETP/FSEL + FMNMX
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
                x[i] = (x[i] < b + i) ? x[i] : a;
                x[i] = fmaxf(x[i], b);
            }
            if (MODE == 6) {                                                              // MUFU + 6 FFMA
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
#pragma unroll
                for (int q = 0; q < 6; ++q) x[i] = fmaf(x[i], a, b);
            }
            if (MODE == 7) {                                                              // F2FP pack
                unsigned u;
                asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(x[i]), "f"(x[(i + 1) & 7]));
                x[i] = __uint_as_float(u) + a;
            }
            if (MODE == 8) x[i] = x[i] + a;                                               // FADD
            if (MODE == 9) {                                                              // integer ISETP + SEL
                int v = __float_as_int(x[i]);
                v = (v < (int)threadIdx.x + i) ? v : 5;
                x[i] = __int_as_float(v + 1);
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int n_instr, int warps_per_smsp) {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    const int threads = warps_per_smsp * 4 * 32;
    k<MODE><<<148, threads>>>(out, cyc, 0.5f, 0.25f);
    k<MODE><<<148, threads>>>(out, cyc, 0.5f, 0.25f);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = (double)h[0] / ((double)ITERS * 8 * warps_per_smsp);
    printf("%-28s warps/SMSP=%d: %.2f cycles per warp-group-of-%d-instr per SMSP (%.2f per instr)\n", name, warps_per_smsp, c,
           n_instr, c / n_instr);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int w : {1, 2, 4}) {
        run<0>("FFMA", 1, w);
        run<1>("MUFU.EX2", 1, w);
        run<2>("FMNMX", 1, w);
        run<3>("FSETP+FSEL", 2, w);
        run<4>("MUFU+2FFMA", 3, w);
        run<5>("MUFU+FSETP+FSEL+FMNMX", 4, w);
        run<6>("MUFU+6FFMA", 7, w);
        run<7>("F2FP+FADD", 2, w);
        run<8>("FADD", 1, w);
        run<9>("ISETP+SEL+IADD", 3, w);
    }
    return 0;
}
