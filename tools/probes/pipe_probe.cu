// pipe_probe.cu -- instruction-throughput probe for the scorer's inner loop (B200, sm_100a).
// For each candidate instruction: 8 independent dependency chains per thread, 256 issues per chain, 32 warps per SM
// -> warp-instructions per cycle per SM (4.0 = one per scheduler per cycle).  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu && ./pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

#define REP 256
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(1024) k_probe(unsigned *out, unsigned seed, long long *cyc)
{
    unsigned r[CHAINS];
    float f[CHAINS];
    float2 g[CHAINS];
    __shared__ signed char sh[34832];
    for (int i = threadIdx.x; i < 34832; i += blockDim.x) sh[i] = (signed char)i;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { r[c] = seed * (c + 1) + threadIdx.x; f[c] = (float)r[c]; g[c] = make_float2(f[c], f[c] + 1.f); }
    const unsigned b = seed | 1u;
    const float fb = 1.0000001f, fc = 0.5f;
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sh);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < REP / 8; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) {
                if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[c]) : "f"(fb), "f"(fc));
                if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(*(unsigned long long *)&g[c]) : "l"(*(const unsigned long long *)&g[(c + 1) & 7]), "l"(*(const unsigned long long *)&g[(c + 2) & 7]));
                if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[c]) : "r"(b), "r"(seed));
                if (OP == 3) asm volatile("mad.hi.u32 %0, %0, 0x11000, %1;" : "+r"(r[c]) : "r"(seed));
                if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x26BB;" : "+r"(r[c]) : "r"(b));
                if (OP == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[c]) : "r"(b), "r"(seed));
                if (OP == 6) asm volatile("{.reg .pred p; .reg .b32 t; and.b32 t, %0, 0xFF80; setp.ne.u32 p, t, 0; @p add.u32 %0, %0, %1;}" : "+r"(r[c]) : "r"(b));
                if (OP == 7) asm volatile("shf.l.wrap.b32 %0, %0, %0, 5;" : "+r"(r[c]));
                if (OP == 8) asm volatile("{.reg .b32 t; shr.u32 t, %0, 4; add.u32 %0, %0, t;}" : "+r"(r[c]));                  // LEA.HI-like
                if (OP == 9) asm volatile("{.reg .b32 a, v; and.b32 a, %0, 0x7fff; add.u32 a, a, %1; ld.shared.s8 v, [a]; add.u32 %0, %0, v;}" : "+r"(r[c]) : "r"(sbase));
                if (OP == 10) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[c]) : "r"(b));
                if (OP == 11) asm volatile("mul.lo.u32 %0, %0, 65536;" : "+r"(r[c]));
                if (OP == 12) asm volatile("{.reg .b32 a, v; mad.hi.u32 a, %0, 0x11000, %1; ld.shared.s8 v, [a]; xor.b32 %0, %0, v;}" : "+r"(r[c]) : "r"(sbase));
                if (OP == 13) asm volatile("mad.wide.u32 %0, %1, 0x11000, %0;" : "+l"(*(unsigned long long *)&g[c]) : "r"(r[c]));
            }
        }
    }
    long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc += r[c] + __float_as_uint(f[c]) + __float_as_uint(g[c].x) + __float_as_uint(g[c].y);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int instr_per_op, unsigned *out, long long *cyc)
{
    const int blocks = 148 * 2;
    k_probe<OP><<<blocks, 1024>>>(out, 12345u, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_probe<OP><<<blocks, 1024>>>(out, 12345u, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[296];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double mean = 0; for (int i = 0; i < blocks; i++) mean += h[i]; mean /= blocks;
    // per SM: 2 blocks x 32 warps, each REP*CHAINS ops
    const double warp_ops = 2.0 * 32 * REP * CHAINS;
    printf("%-34s %8.0f clock64 ticks, %7.1f us by events (= %8.0f cycles at 1965 MHz)  %6.3f ops/cycle/SM by events  (%d SASS instr per op)  %s\n", name, mean,
           ms * 1e3, ms * 1e-3 * 1965e6, warp_ops / (ms * 1e-3 * 1965e6), instr_per_op, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    unsigned *out; long long *cyc;
    cudaMalloc(&out, 4 * 296 * 1024); cudaMalloc(&cyc, 8 * 296);
    run<0>("FFMA", 1, out, cyc);
    run<1>("FFMA2 (fma.rn.f32x2)", 1, out, cyc);
    run<2>("IMAD (mad.lo)", 1, out, cyc);
    run<3>("IMAD.HI (mad.hi.u32 imm)", 1, out, cyc);
    run<13>("IMAD.WIDE (mad.wide.u32)", 1, out, cyc);
    run<11>("IMAD.SHL (mul.lo by 65536)", 1, out, cyc);
    run<4>("PRMT", 1, out, cyc);
    run<5>("LOP3", 1, out, cyc);
    run<6>("LOP3.P + predicated IADD", 2, out, cyc);
    run<7>("SHF", 1, out, cyc);
    run<8>("shr+add (LEA.HI)", 1, out, cyc);
    run<10>("IADD", 1, out, cyc);
    run<9>("LOP+IADD+LDS.S8+IADD", 4, out, cyc);
    run<12>("IMAD.HI+LDS.S8+LOP", 3, out, cyc);
    return 0;
}
