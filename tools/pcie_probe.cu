// H2D / D2H rate of differently allocated pinned host buffers (build: nvcc -O2 -o build/pcie_probe tools/pcie_probe.cu)
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
static float time_copy(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t s) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) cudaMemcpyAsync(dst, src, n, k, s);
    cudaStreamSynchronize(s);
    cudaEventRecord(e0, s);
    for (int i = 0; i < 10; ++i) cudaMemcpyAsync(dst, src, n, k, s);
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / 10;
}
int main() {
    const size_t n = 30932992;
    void *d; cudaMalloc(&d, n);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (int trial = 0; trial < 3; ++trial) {
        void *h;
        cudaHostAlloc(&h, n, cudaHostAllocDefault); memset(h, 1, n);
        printf("trial %d hostalloc default : h2d %.1f GB/s  d2h %.1f GB/s\n", trial, n / 1e6 / time_copy(d, h, n, cudaMemcpyHostToDevice, s), n / 1e6 / time_copy(h, d, n, cudaMemcpyDeviceToHost, s));
        cudaFreeHost(h);
        cudaHostAlloc(&h, n, cudaHostAllocWriteCombined); memset(h, 1, n);
        printf("trial %d hostalloc WC      : h2d %.1f GB/s  d2h %.1f GB/s\n", trial, n / 1e6 / time_copy(d, h, n, cudaMemcpyHostToDevice, s), n / 1e6 / time_copy(h, d, n, cudaMemcpyDeviceToHost, s));
        cudaFreeHost(h);
        const size_t huge = 2u << 20, nn = (n + huge - 1) / huge * huge;
        void *m = mmap(nullptr, nn + huge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        void *al = (void *)(((size_t)m + huge - 1) / huge * huge);
        madvise(al, nn, MADV_HUGEPAGE);
        memset(al, 1, nn);
        cudaError_t e = cudaHostRegister(al, nn, cudaHostRegisterDefault);
        printf("trial %d mmap THP register (%s): h2d %.1f GB/s  d2h %.1f GB/s\n", trial, cudaGetErrorString(e), n / 1e6 / time_copy(d, al, n, cudaMemcpyHostToDevice, s), n / 1e6 / time_copy(al, d, n, cudaMemcpyDeviceToHost, s));
        cudaHostUnregister(al);
        munmap(m, nn + huge);
        void *mh = mmap(nullptr, nn, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_HUGETLB, -1, 0);
        if (mh != MAP_FAILED) {
            memset(mh, 1, nn);
            e = cudaHostRegister(mh, nn, cudaHostRegisterDefault);
            printf("trial %d MAP_HUGETLB register (%s): h2d %.1f GB/s\n", trial, cudaGetErrorString(e), n / 1e6 / time_copy(d, mh, n, cudaMemcpyHostToDevice, s));
            cudaHostUnregister(mh); munmap(mh, nn);
        } else printf("trial %d MAP_HUGETLB unavailable\n", trial);
        // chunked: 8 copies of n/8
        cudaHostAlloc(&h, n, cudaHostAllocDefault); memset(h, 1, n);
        {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0, s);
            for (int r = 0; r < 10; ++r) for (int c = 0; c < 8; ++c) cudaMemcpyAsync((char *)d + c * (n / 8), (char *)h + c * (n / 8), n / 8, cudaMemcpyHostToDevice, s);
            cudaEventRecord(e1, s); cudaStreamSynchronize(s);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("trial %d hostalloc default, 8 chunks: h2d %.1f GB/s\n", trial, n / 1e6 / (ms / 10));
        }
        cudaFreeHost(h);
    }
    printf("THP: "); fflush(stdout); system("cat /sys/kernel/mm/transparent_hugepage/enabled");
    return 0;
}
