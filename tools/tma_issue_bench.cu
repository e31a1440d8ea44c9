// What does it cost the issuing thread to start a TMA bulk copy (cp.async.bulk global -> shared, UBLKCP)?  One CTA; the
// producer issues `n` copies of `bytes` each into a ring of shared-memory buffers, every copy on its own mbarrier, and
// only then waits for them.  Reported: cycles per issue (clock64 around the issue loop only) and cycles until all arrived.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/tma_issue_bench.cu -o build/tma_issue_bench
#include <cstdio>
#include "../gprotation_b200/csrc/gp_kernel.cuh"
using namespace gprot;

constexpr int NBUF = 8;

template <int MODE>   // 0: if (tid == 0) in a 256-thread CTA (the kernels' form)   1: one elected lane of warp 0, warp-uniform branch
__global__ void __launch_bounds__(256, 1) k_issue(const double* src, long long* out, int n, int bytes, int stride_bytes) {
    extern __shared__ __align__(128) unsigned char raw[];
    __shared__ unsigned long long bar[NBUF];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NBUF; i++) mbar_init(&bar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long t_issue = 0, t_all = 0;
    for (int rep = 0; rep < 6; rep++) {
        bool me;
        if (MODE == 0) me = tid == 0;
        else {
            uint32_t el = 0;
            if (tid < 32) asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(el));
            me = el != 0;
        }
        const long long t0 = clock64();
        if (me) {
            for (int i = 0; i < n; i++) {
                mbar_expect_tx(&bar[i % NBUF], bytes);
                tma_load(raw + (size_t)(i % NBUF) * 16384, reinterpret_cast<const char*>(src) + ((size_t)rep * n + i) * stride_bytes, bytes, &bar[i % NBUF]);
            }
        }
        const long long t1 = clock64();
        if (tid == 0 || MODE == 1) {
            if (tid < 32) for (int i = 0; i < n; i++) mbar_wait(&bar[i % NBUF], rep & 1);
        }
        const long long t2 = clock64();
        __syncthreads();
        if (me && rep > 0) { t_issue += t1 - t0; t_all += t2 - t0; }
        if (me && rep == 5) { out[0] = t_issue / 5; out[1] = t_all / 5; }
    }
}

int main() {
    double* src; long long* out;
    cudaMalloc(&src, (size_t)64 << 20); cudaMemset(src, 0, (size_t)64 << 20);
    cudaMalloc(&out, 64);
    cudaFuncSetAttribute(k_issue<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, NBUF * 16384);
    cudaFuncSetAttribute(k_issue<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, NBUF * 16384);
    for (int mode = 0; mode < 2; mode++)
        for (int bytes : {1024, 8192, 16384})
            for (int n : {1, 2, 4, 8}) {
                if (mode == 0) k_issue<0><<<1, 256, NBUF * 16384>>>(src, out, n, bytes, 16384);
                else k_issue<1><<<1, 256, NBUF * 16384>>>(src, out, n, bytes, 16384);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2];
                cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                printf("%s, %5d bytes x %d: %6lld cycles in the issue loop (%5lld per copy), all arrived after %6lld  %s\n",
                       mode ? "elected lane of warp 0" : "if (tid == 0)         ", bytes, n, h[0], h[0] / n, h[1], e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
