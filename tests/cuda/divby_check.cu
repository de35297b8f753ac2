// Checks cb::DivBy (csrc/divby.cuh) against the compiler's a / b, bit for bit, on ~2e10 operand pairs: fit_line()-like
// magnitudes, random mantissas over a wide exponent range, and completely random bit patterns (denormals, infinities, NaN).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "divby.cuh"

__global__ void k(unsigned long long seed, unsigned long long *bad, double *ex)
{
    unsigned long long s = seed + (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull;
    for (int it = 0; it < 4096; it++) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        const int mode = it & 3;
        double b, a;
        if (mode == 0) b = __longlong_as_double((long long)((s & 0x000fffffffffffffull) | ((0x3ffull + (s >> 60)) << 52)));
        else if (mode == 1) b = (double)((s >> 20) & 0xffffff) * 0.25 + 1.0;
        else b = __longlong_as_double((long long)(s & 0x7fffffffffffffffull));
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        if (mode == 0) a = __longlong_as_double((long long)((s & 0x000fffffffffffffull) | ((0x3f0ull + (s >> 58)) << 52)));
        else if (mode == 1) a = (double)(long long)(s >> 24) * 0.125;
        else a = __longlong_as_double((long long)s);
        if (b != b || a != a) continue;
        const cb::DivBy d(b);
        const double q1 = d(a), q2 = a / b;
        if (__double_as_longlong(q1) != __double_as_longlong(q2) && !(q1 != q1 && q2 != q2))
            if (atomicAdd(bad, 1ull) == 0) { ex[0] = a; ex[1] = b; ex[2] = q1; ex[3] = q2; }
    }
}

int main()
{
    unsigned long long *bad;
    double *ex;
    if (cudaMallocManaged(&bad, 8) != cudaSuccess || cudaMallocManaged(&ex, 32) != cudaSuccess) { printf("no device\n"); return 2; }
    *bad = 0;
    for (int r = 0; r < 8; r++) k<<<148 * 16, 256>>>(0x1234567ull + r * 7919ull, bad, ex);
    const cudaError_t e = cudaDeviceSynchronize();
    printf("mismatches %llu of %.3g  first: a=%a b=%a DivBy=%a a/b=%a  (%s)\n", *bad, 8.0 * 148 * 16 * 256 * 4096, ex[0], ex[1], ex[2], ex[3],
           cudaGetErrorString(e));
    return (*bad != 0 || e != cudaSuccess) ? 1 : 0;
}
