// microbench.cu -- per-SM instruction throughput of the integer ops the kernels are built from, measured on
// the target GPU (B200). Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu
// Output: ops/clk/SM for each op class (dependent chains broken into 8 independent accumulators per thread).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define NACC 8

template <int OP>
__global__ void k(uint32_t *out, uint32_t a0, uint32_t b0, long long *cyc) {
    uint32_t acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = a0 + threadIdx.x * 7 + i;
    float facc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) facc[i] = 1.0f + (float)(threadIdx.x + i);
    uint32_t b = b0 + threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (OP == 0) acc[i] = acc[i] * b + 12345u;                                     // IMAD
            if (OP == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[i]) : "r"(b));        // IADD (may become IMAD.IADD)
            if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(acc[i]) : "r"(b), "r"(a0));   // LOP3
            if (OP == 3) asm volatile("shr.s32 %0, %0, 3;" : "+r"(acc[i]));                  // SHF
            if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(acc[i]) : "r"(b)); // PRMT
            if (OP == 5) asm volatile("dp4a.s32.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(b), "r"(a0)); // IDP.4A
            if (OP == 6) asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(b), "r"(a0));   // IMAD.HI
            if (OP == 7) acc[i] = (uint32_t)((int32_t)acc[i] >> 16) + b;                     // LEA.HI.SX32 candidate
            if (OP == 8) facc[i] = facc[i] * 1.0001f + 0.5f;                                  // FFMA
            if (OP == 9) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(facc[i]));        // MUFU.RSQ
            if (OP == 10) { float f = __int2float_rn((int)acc[i]); acc[i] = __float_as_uint(f) ^ b; }   // I2F
            if (OP == 11) acc[i] = __usad(acc[i], b, acc[i]);                                 // VABSDIFF
            if (OP == 12) asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(b), "r"(a0)); // IDP.2A
            if (OP == 13) { acc[i] = acc[i] * b + 12345u; facc[i] = facc[i] * 1.0001f + 0.5f; } // IMAD + FFMA mix
            if (OP == 14) { acc[i] = acc[i] * b + 12345u; asm volatile("shr.s32 %0, %0, 1;" : "+r"(acc[(i + 4) & 7])); } // IMAD + SHF mix
            if (OP == 15) asm volatile("vabsdiff.u32.u32.u32.add %0, %0, %1, %0;" : "+r"(acc[i]) : "r"(b));
            if (OP == 16) { unsigned long long w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(acc[i]), "r"(b)); acc[i] = (uint32_t)(w >> 32) + (uint32_t)w; }  // IMAD.WIDE.U32 + IADD
            if (OP == 17) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(acc[i]) : "r"(b));                  // IMAD.HI.U32
            if (OP == 18) { unsigned long long w; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(acc[i]), "r"(b)); acc[i] = (uint32_t)(w >> 37); }   // udiv-by-constant via WIDE + SHF
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i] + __float_as_uint(facc[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char *name, int ops_per_iter) {
    uint32_t *out; long long *cyc, h;
    const int threads = 1024, blocks = 148;   // one full CTA per SM
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, 8);
    k<OP><<<blocks, threads>>>(out, 3, 5, cyc);
    k<OP><<<blocks, threads>>>(out, 3, 5, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double ops = (double)ITER * NACC * threads * ops_per_iter;
    printf("%-28s %8.1f thread-ops/clk/SM   (%lld cycles)\n", name, ops / (double)h, h);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("IMAD", 1); run<1>("IADD", 1); run<2>("LOP3", 1); run<3>("SHF", 1); run<4>("PRMT", 1);
    run<5>("IDP.4A (dp4a)", 1); run<12>("IDP.2A (dp2a)", 1); run<6>("IMAD.HI", 1); run<7>("a>>16 + b (LEA.HI?)", 1);
    run<8>("FFMA", 1); run<9>("MUFU.RSQ", 1); run<10>("I2F + LOP", 2); run<11>("usad", 1); run<15>("vabsdiff.add", 1);
    run<13>("IMAD + FFMA", 2); run<14>("IMAD + SHF", 2);
    run<16>("IMAD.WIDE.U32 + IADD", 2); run<17>("IMAD.HI.U32 (mul.hi)", 1); run<18>("IMAD.WIDE.U32 + SHF", 2);
    return 0;
}
