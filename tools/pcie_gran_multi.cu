// Micro-benchmark behind the e2e scaling curve: N GPUs of one box read pinned host memory AT THE SAME
// TIME - scattered 4-byte reads at the detector's texel pitch (60 B on a 1080p RGBA row: one 64-byte line per read) and at a
// contiguous pitch, plus bulk H2D copies - and each reports its own rate.  If the per-GPU rate of the scattered reads falls
// as N grows while each GPU has its own x16 link, the bound is the host side (root complex / memory system), not the links.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xcompiler -pthread -o /tmp/pcie_gran_multi tools/pcie_gran_multi.cu
//   /tmp/pcie_gran_multi            # runs N = 1, 2, 4, 8 (as many as the box has)
#include <cuda_runtime.h>
#include <pthread.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void __launch_bounds__(256) reads(const unsigned char *base, long long stride, long long n, unsigned *out) {
    const long long step = (long long)gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n; i0 += 4 * step) {
        unsigned v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long i = i0 + u * step;
            v[u] = i < n ? __ldg(reinterpret_cast<const unsigned *>(base + i * stride)) : 0u;
        }
        acc += v[0] + v[1] + v[2] + v[3];
    }
    if (acc == 0xdeadbeefu) out[0] = acc;
}

struct Worker {
    int dev, n_active;
    unsigned char *h;        // this GPU's own pinned buffer
    size_t bytes;
    pthread_barrier_t *bar;
    double ns60, ns4, gbs_copy;
};

static void *work(void *arg) {
    Worker *w = static_cast<Worker *>(arg);
    cudaSetDevice(w->dev);
    unsigned *out;
    cudaMalloc(&out, 4);
    unsigned char *d;
    const size_t copy_bytes = 512ull << 20;
    cudaMalloc(&d, copy_bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    reads<<<148, 256>>>(w->h, 60, 1024, out);
    cudaDeviceSynchronize();
    const long long nn = 8ll << 20;                 // 8 Mi reads: 60 B pitch spans 480 MiB, 4 B pitch 32 MiB
    float ms;
    for (int pass = 0; pass < 3; pass++) {
        pthread_barrier_wait(w->bar);               // all active GPUs start together
        cudaEventRecord(a);
        if (pass == 0) reads<<<148, 256>>>(w->h, 60, nn, out);
        else if (pass == 1) reads<<<148, 256>>>(w->h + (512ull << 20), 4, nn, out);
        else cudaMemcpyAsync(d, w->h, copy_bytes, cudaMemcpyHostToDevice);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        if (pass == 0) w->ns60 = ms * 1e6 / (double)nn;
        else if (pass == 1) w->ns4 = ms * 1e6 / (double)nn;
        else w->gbs_copy = copy_bytes / 1e9 / (ms / 1e3);
        pthread_barrier_wait(w->bar);
    }
    cudaFree(d);
    cudaFree(out);
    return nullptr;
}

int main() {
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (ndev < 1) return printf("no device\n"), 1;
    const size_t bytes = 1ull << 30;
    std::vector<unsigned char *> bufs(ndev, nullptr);
    for (int i = 0; i < ndev; i++) {
        cudaSetDevice(i);
        if (cudaHostAlloc(&bufs[i], bytes, cudaHostAllocPortable) != cudaSuccess) return printf("pinned alloc failed\n"), 1;
        for (size_t k = 0; k < bytes; k += 4096) bufs[i][k] = (unsigned char)k;
    }
    printf("%-4s %-5s %14s %16s %14s %16s %12s\n", "N", "GPU", "ns / 60B read", "lines GB/s (64B)", "ns / 4B read", "contig GB/s", "H2D GB/s");
    for (int n = 1; n <= ndev; n *= 2) {
        pthread_barrier_t bar;
        pthread_barrier_init(&bar, nullptr, n);
        std::vector<Worker> ws(n);
        std::vector<pthread_t> th(n);
        for (int i = 0; i < n; i++) {
            ws[i] = Worker{i, n, bufs[i], bytes, &bar, 0, 0, 0};
            pthread_create(&th[i], nullptr, work, &ws[i]);
        }
        double sum_lines = 0, sum_copy = 0;
        for (int i = 0; i < n; i++) pthread_join(th[i], nullptr);
        for (int i = 0; i < n; i++) {
            const double lines = 64.0 / ws[i].ns60, contig = 4.0 / ws[i].ns4;
            sum_lines += lines, sum_copy += ws[i].gbs_copy;
            printf("%-4d %-5d %14.3f %16.1f %14.3f %16.1f %12.1f\n", n, i, ws[i].ns60, lines, ws[i].ns4, contig, ws[i].gbs_copy);
        }
        printf("%-4d %-5s %14s %16.1f %14s %16s %12.1f   <- box totals\n", n, "sum", "", sum_lines, "", "", sum_copy);
        pthread_barrier_destroy(&bar);
    }
    return 0;
}
