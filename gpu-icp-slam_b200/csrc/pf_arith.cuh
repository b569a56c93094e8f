// pf_arith.cuh -- scalar device arithmetic of the particle filter.
//
// Every floating-point step is an explicit round-to-nearest IEEE-754 binary32 intrinsic
// (__fmaf_rn / __fmul_rn / __fadd_rn / __fdiv_rn / __fsqrt_rn), so nvcc can neither contract nor
// reassociate it and the results are reproducible on any IEEE machine.
//
// Reference semantics (paths relative to michaelwillett/GPU-ICP-SLAM):
//   pf_hash            src/kernel.cu:89-97   utilhash
//   pf_seed            src/kernel.cu:99-102  makeSeededRandomEngine (int shifts wrap)
//   pf_minstd_*        thrust::minstd_rand (default_random_engine) seed / step
//   pf_normal          thrust normal_distribution_nvcc::sample (normal_distribution_base.h:50-80),
//                      with CUDA's MUFU-based erfcinvf replaced by an IEEE-only erfcinv
//   pf_lidar_angle     src/kernel.cu:42      LIDAR_ANGLE(i)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace pf {

__device__ __forceinline__ uint32_t pf_hash(uint32_t a)
{
    a = (a + 0x7ed55d16u) + (a << 12);
    a = (a ^ 0xc761c23cu) ^ (a >> 19);
    a = (a + 0x165667b1u) + (a << 5);
    a = (a + 0xd3a2646cu) ^ (a << 9);
    a = (a + 0xfd7046c5u) + (a << 3);
    a = (a ^ 0xb55a4f09u) ^ (a >> 16);
    return a;
}

__device__ __forceinline__ uint32_t pf_seed(int iter, int index, int depth)
{
    uint32_t k = (1u << 31) | ((uint32_t)depth << 22) | (uint32_t)iter;
    return pf_hash(k) ^ pf_hash((uint32_t)index);
}

__device__ __forceinline__ uint32_t pf_minstd_seed(uint32_t s)
{
    uint32_t x = s % 2147483647u;
    return x == 0u ? 1u : x;
}

__device__ __forceinline__ uint32_t pf_minstd_next(uint32_t &state)
{
    state = (uint32_t)(((uint64_t)state * 48271ull) % 2147483647ull);
    return state;
}

// natural log of a normal positive float; Cephes logf scheme, fixed operation order
__device__ __forceinline__ float pf_logf(float x)
{
    uint32_t ix = __float_as_uint(x);
    int e = (int)(ix >> 23) - 126;
    float m = __uint_as_float((ix & 0x007fffffu) | 0x3f000000u);
    float f;
    if (m < 0.707106769084930419921875f) { e -= 1; f = __fsub_rn(__fadd_rn(m, m), 1.0f); }
    else f = __fsub_rn(m, 1.0f);
    float z = __fmul_rn(f, f);
    float p = 7.0376836292E-2f;
    p = __fmaf_rn(p, f, -1.1514610310E-1f);
    p = __fmaf_rn(p, f, 1.1676998740E-1f);
    p = __fmaf_rn(p, f, -1.2420140846E-1f);
    p = __fmaf_rn(p, f, 1.4249322787E-1f);
    p = __fmaf_rn(p, f, -1.6668057665E-1f);
    p = __fmaf_rn(p, f, 2.0000714765E-1f);
    p = __fmaf_rn(p, f, -2.4999993993E-1f);
    p = __fmaf_rn(p, f, 3.3333331174E-1f);
    float y = __fmul_rn(__fmul_rn(p, f), z);
    float fe = (float)e;
    y = __fmaf_rn(fe, -2.12194440e-4f, y);
    y = __fmaf_rn(-0.5f, z, y);
    float r = __fadd_rn(f, y);
    r = __fmaf_rn(fe, 0.693359375f, r);
    return r;
}

// erfcinv on (0,2): Giles' single-precision erfinv polynomials on w = -log(y(2-y)), times (1-y)
__device__ __forceinline__ float pf_erfcinvf(float y)
{
    float t = __fmul_rn(y, __fsub_rn(2.0f, y));
    float w = __fsub_rn(0.0f, pf_logf(t));
    float p;
    if (w < 5.0f) {
        w = __fsub_rn(w, 2.5f);
        p = 2.81022636e-08f;
        p = __fmaf_rn(p, w, 3.43273939e-07f);
        p = __fmaf_rn(p, w, -3.5233877e-06f);
        p = __fmaf_rn(p, w, -4.39150654e-06f);
        p = __fmaf_rn(p, w, 0.00021858087f);
        p = __fmaf_rn(p, w, -0.00125372503f);
        p = __fmaf_rn(p, w, -0.00417768164f);
        p = __fmaf_rn(p, w, 0.246640727f);
        p = __fmaf_rn(p, w, 1.50140941f);
    } else {
        w = __fsub_rn(__fsqrt_rn(w), 3.0f);
        p = -0.000200214257f;
        p = __fmaf_rn(p, w, 0.000100950558f);
        p = __fmaf_rn(p, w, 0.00134934322f);
        p = __fmaf_rn(p, w, -0.00367342844f);
        p = __fmaf_rn(p, w, 0.00573950773f);
        p = __fmaf_rn(p, w, -0.0076224613f);
        p = __fmaf_rn(p, w, 0.00943887047f);
        p = __fmaf_rn(p, w, 1.00167406f);
        p = __fmaf_rn(p, w, 2.83297682f);
    }
    return __fmul_rn(p, __fsub_rn(1.0f, y));
}

// one normal variate with mean 0; S1 = float(1/2147483645.) = 2^-31, S2 = 2^-32, S3 = -+sqrt(2)
__device__ __forceinline__ float pf_normal(uint32_t &state, float stddev)
{
    uint32_t u = pf_minstd_next(state) - 1u;
    float s3 = -1.41421354f;
    if (u > 1073741822u) { u = 2147483645u - u; s3 = 1.41421354f; }
    float p = __fmaf_rn((float)u, 4.656612873077392578125e-10f, 2.3283064365386962890625e-10f);
    float k = __fmul_rn(stddev, s3);
    return __fmul_rn(k, pf_erfcinvf(__fadd_rn(p, p)));
}

__device__ __forceinline__ float pf_lidar_angle(int i)
{
    float a = __fadd_rn(-135.0f, __fmul_rn((float)i, 0.25f));
    a = __fmul_rn(a, 3.1415926535897932384626422832795028841971f);
    return __fdiv_rn(a, 180.0f);
}

}  // namespace pf
