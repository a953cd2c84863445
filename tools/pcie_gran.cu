// Micro-benchmark: what does ONE scattered 4-byte read of pinned host memory cost across PCIe, as a function of the
// distance between reads and of the load flavour?  (the pinned-host ingest of bench `e2e` is
// bound by exactly these reads.)   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pcie_gran tools/pcie_gran.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

template <int MODE>
__device__ __forceinline__ unsigned load(const unsigned *p) {
    unsigned v;
    if (MODE == 0) v = __ldg(p);
    else if (MODE == 1) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (MODE == 2) asm volatile("ld.global.cv.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (MODE == 3) asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else if (MODE == 4) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    else asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(256) reads(const unsigned char *base, long long stride, long long n, unsigned *out) {
    const long long step = (long long)gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * step) {
        unsigned v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = i0 + u * step;
            v[u] = i < n ? load<MODE>(reinterpret_cast<const unsigned *>(base + i * stride)) : 0u;
        }
        acc += v[0] + v[1] + v[2] + v[3];
    }
    if (acc == 0xdeadbeefu) out[0] = acc;
}

template <int MODE>
float run(const unsigned char *d, long long stride, long long n, unsigned *out) {
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    reads<MODE><<<148, 256>>>(d, stride, 1024, out);   // warm-up (code load only: no lines of the timed pass cached)
    cudaEventRecord(a);
    reads<MODE><<<148, 256>>>(d, stride, n, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    const size_t bytes = 2ull << 30;
    unsigned char *h = nullptr;
    if (cudaHostAlloc(&h, bytes, cudaHostAllocMapped) != cudaSuccess) return printf("pinned alloc failed\n"), 1;
    for (size_t i = 0; i < bytes; i += 4096) h[i] = (unsigned char)i;
    unsigned *out;
    cudaMalloc(&out, 4);
    const char *names[] = {"ldg(nc)", "volatile", "cv", "cs", "nc.no_alloc", "relaxed.sys"};
    const long long strides[] = {4, 16, 32, 60, 64, 128, 256, 512};
    printf("%-12s", "stride B:");
    for (long long s : strides) printf("%10lld", s);
    printf("   (ns per read, M reads/s in brackets at 60 B)\n");
    for (int m = 0; m < 6; m++) {
        printf("%-12s", names[m]);
        for (long long s : strides) {
            const long long nn = 4ll << 20;   // 4 Mi reads, span <= 2 GiB
            // short spans: a fresh 256 MiB window per flavour, so nothing is left in L2 from the previous one
            const unsigned char *hb = s <= 64 ? h + (size_t)m * (256ull << 20) : h;
            float ms = 0;
            switch (m) {
                case 0: ms = run<0>(hb, s, nn, out); break;
                case 1: ms = run<1>(hb, s, nn, out); break;
                case 2: ms = run<2>(hb, s, nn, out); break;
                case 3: ms = run<3>(hb, s, nn, out); break;
                case 4: ms = run<4>(hb, s, nn, out); break;
                default: ms = run<5>(hb, s, nn, out); break;
            }
            printf("%10.2f", ms * 1e6 / (double)nn);
        }
        printf("\n");
    }
    // reference point: bulk copy of 1 GiB
    unsigned char *d;
    cudaMalloc(&d, 1ull << 30);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    cudaMemcpy(d, h, 1ull << 30, cudaMemcpyHostToDevice);
    cudaEventRecord(a);
    cudaMemcpyAsync(d, h, 1ull << 30, cudaMemcpyHostToDevice);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    printf("cudaMemcpy H2D 1 GiB: %.2f ms = %.1f GB/s\n", ms, 1.073741824 / ms * 1e3);
    return 0;
}
