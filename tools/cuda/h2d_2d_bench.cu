// h2d_2d_bench.cu -- does a strided host->device copy (even rows only) run at the rate of a contiguous one?  And how fast can a kernel
// pull scattered row pieces straight from pinned host memory (zero-copy)?   nvcc -O3 -o tools/cuda/h2d_2d_bench tools/cuda/h2d_2d_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// one warp per (box, odd row): copies `bw` bytes of a row piece from mapped host memory into the device image
__global__ void fetch_boxes_kernel(const unsigned char *__restrict__ host_img, unsigned char *__restrict__ dev_img, int stride, size_t frame_stride,
                                   int nboxes_per_frame, int bw, int bh, int W, int H, int nframes)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int rows_per_box = bh / 2;
    const long long total = (long long)nframes * nboxes_per_frame * rows_per_box;
    if (warp >= total) return;
    const int r = (int)(warp % rows_per_box);
    const long long q = warp / rows_per_box;
    const int box = (int)(q % nboxes_per_frame), f = (int)(q / nboxes_per_frame);
    const int x0 = ((box * 197 + f * 31) % (W - bw)) & ~15, y0 = ((box * 113 + f * 17) % (H - bh)) & ~1;
    const size_t off = (size_t)f * frame_stride + (size_t)(y0 + 2 * r + 1) * stride + x0;
    for (int k = lane * 16; k < bw; k += 512)
        *reinterpret_cast<uint4 *>(dev_img + off + k) = *reinterpret_cast<const uint4 *>(host_img + off + k);
}

int main(int argc, char **argv)
{
    const int W = 1280, H = 720, B = argc > 1 ? atoi(argv[1]) : 256;
    const size_t fs = (size_t)W * H, bytes = fs * B;
    unsigned char *h, *d;
    CHECK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
    CHECK(cudaMalloc(&d, bytes));
    for (size_t i = 0; i < bytes; i += 4096) h[i] = (unsigned char)i;
    cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    auto timeit = [&](const char *name, double moved, auto fn) {
        fn(); CHECK(cudaDeviceSynchronize());
        float best = 1e9f, sum = 0;
        for (int i = 0; i < 8; i++) {
            CHECK(cudaEventRecord(e0)); fn(); CHECK(cudaEventRecord(e1)); CHECK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); sum += ms;
        }
        printf("%-58s avg %.3f ms  best %.3f ms  %.1f GB/s\n", name, sum / 8, best, moved / (sum / 8) * 1e-6);
    };
    timeit("contiguous cudaMemcpyAsync, whole frames", (double)bytes, [&] { CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice)); });
    timeit("contiguous cudaMemcpyAsync, half the bytes", (double)bytes / 2, [&] { CHECK(cudaMemcpyAsync(d, h, bytes / 2, cudaMemcpyHostToDevice)); });
    timeit("cudaMemcpy2DAsync, even rows (1280 B of every 2560)", (double)bytes / 2,
           [&] { CHECK(cudaMemcpy2DAsync(d, 2 * W, h, 2 * W, W, (size_t)B * H / 2, cudaMemcpyHostToDevice)); });
    unsigned char *hd = nullptr;
    CHECK(cudaHostGetDevicePointer((void **)&hd, h, 0));
    for (int nb : {8, 16, 32}) {
        const int bw = 160, bh = 160;
        const long long warps = (long long)B * nb * (bh / 2);
        char nm[96]; snprintf(nm, 96, "zero-copy fetch kernel: %d boxes of %dx%d per frame (odd rows)", nb, bw, bh);
        timeit(nm, (double)B * nb * bw * (bh / 2), [&] { fetch_boxes_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256>>>(hd, d, W, fs, nb, bw, bh, W, H, B); });
    }
    return 0;
}
