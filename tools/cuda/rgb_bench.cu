// rgb_bench.cu -- micro-benchmark of the packed-RGB -> gray kernels (round 2 experiments; not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --fmad=false -lineinfo -o tools/cuda/rgb_bench tools/cuda/rgb_bench.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../chalkydri_b200/csrc/threshold.cuh"
using namespace cb;
#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// plain streaming kernel with the same traffic: three 128-bit loads, one 128-bit store per lane (no arithmetic to speak of)
__global__ void __launch_bounds__(256) plain_rgb_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t nout)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nout; i += (size_t)gridDim.x * blockDim.x) {
        const size_t base = (i / 32) * 96 + (i % 32);
        const uint4 a = ldg_stream(in + base), b = ldg_stream(in + base + 32), c = ldg_stream(in + base + 64);
        out[i] = make_uint4(a.x ^ b.x ^ c.x, a.y ^ b.y ^ c.y, a.z ^ b.z ^ c.z, a.w ^ b.w ^ c.w);
    }
}

int main(int argc, char **argv)
{
    const int W = argc > 1 ? atoi(argv[1]) : 1456, H = argc > 2 ? atoi(argv[2]) : 1088, B = argc > 3 ? atoi(argv[3]) : 256;
    const size_t npix = (size_t)W * H, nin = npix * 3 * B, nout = npix * B;
    uint8_t *d_in, *d_out, *d_ref;
    CHECK(cudaMalloc(&d_in, nin + 64)); CHECK(cudaMalloc(&d_out, nout + 64)); CHECK(cudaMalloc(&d_ref, nout + 64));
    std::vector<uint8_t> h(npix * 3 * 4);
    uint32_t s = 1;
    for (auto &v : h) { s = s * 1664525u + 1013904223u; v = (uint8_t)(s >> 24); }
    for (int b = 0; b < B; b++) CHECK(cudaMemcpy(d_in + (size_t)b * npix * 3, h.data() + (size_t)(b % 4) * npix * 3, npix * 3, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    const double bytes = 4.0 * npix * B;
    auto timeit = [&](const char *name, auto launch, bool check) {
        for (int i = 0; i < 2; i++) launch();
        CHECK(cudaDeviceSynchronize());
        float sum = 0, best = 1e9f;
        for (int i = 0; i < 10; i++) {
            CHECK(cudaEventRecord(e0)); launch(); CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); sum += ms; best = std::min(best, ms);
        }
        const char *par = "";
        if (check) {
            std::vector<uint8_t> a(nout), b(nout);
            CHECK(cudaMemcpy(a.data(), d_out, nout, cudaMemcpyDeviceToHost)); CHECK(cudaMemcpy(b.data(), d_ref, nout, cudaMemcpyDeviceToHost));
            par = memcmp(a.data(), b.data(), nout) == 0 ? " same" : " DIFFERENT";
        }
        printf("%-44s avg %.4f ms (%.0f GB/s, %.3f)  best %.4f ms%s\n", name, sum / 10, bytes / (sum / 10) * 1e-6, bytes / (sum / 10) * 1e-6 / 6533.2, best, par);
    };
    printf("%dx%d x %d frames packed RGB, %.1f MB algorithmic per launch\n", W, H, B, bytes * 1e-6);
    uint8_t *d_tmp;
    CHECK(cudaMalloc(&d_tmp, (size_t)(bytes / 2)));
    timeit("cudaMemcpy D2D of the same traffic", [&] { CHECK(cudaMemcpyAsync(d_tmp, d_in, (size_t)(bytes / 2), cudaMemcpyDeviceToDevice)); }, false);
    CHECK(cudaFree(d_tmp));
    for (int mult : {4, 8, 16}) {
        char nm[64]; snprintf(nm, 64, "plain LDG.128 x3 / STG.128 stream, 148x%d CTAs", mult);
        timeit(nm, [&] { plain_rgb_kernel<<<148 * mult, 256>>>((const uint4 *)d_in, (uint4 *)d_out, nout / 16); }, false);
    }
    const dim3 grid_rgb((unsigned)((npix + RGB_PX_PER_BLOCK - 1) / RGB_PX_PER_BLOCK), (unsigned)B);
    timeit("rgb_to_gray_kernel (round 1: CTA refill)", [&] { rgb_to_gray_kernel<<<grid_rgb, RGB_THREADS>>>(d_in, d_out, npix, npix * 3, npix); }, false);
    CHECK(cudaMemcpy(d_ref, d_out, nout, cudaMemcpyDeviceToDevice));
    for (int mult : {2, 4, 6, 8, 12, 16}) {
        char nm[64]; snprintf(nm, 64, "rgb_to_gray_ring_kernel, 148x%d CTAs", mult);
        CHECK(cudaMemset(d_out, 0, nout));
        timeit(nm, [&] { rgb_to_gray_ring_kernel<<<148 * mult, RGB_RING_WARPS * 32>>>(d_in, d_out, npix, npix * 3, npix, B); }, true);
    }
    return 0;
}
