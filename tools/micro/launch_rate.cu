// Micro-benchmark: aggregate kernel-launch throughput from T host threads (one stream each) on one GPU.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <thread>
#include <vector>
__global__ void tiny(int* p) { if (p && threadIdx.x == 9999) *p = 1; }
int main() {
    cudaFree(0);
    for (int T : {1, 2, 4, 8}) {
        const int n = 20000;
        std::vector<cudaStream_t> st(T);
        for (auto& s : st) cudaStreamCreate(&s);
        tiny<<<1, 32>>>(nullptr); cudaDeviceSynchronize();
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back([&, t] { for (int i = 0; i < n; ++i) tiny<<<1, 32, 0, st[t]>>>(nullptr); });
        for (auto& x : th) x.join();
        auto t1 = std::chrono::steady_clock::now();
        cudaDeviceSynchronize();
        auto t2 = std::chrono::steady_clock::now();
        double enq = std::chrono::duration<double>(t1 - t0).count(), all = std::chrono::duration<double>(t2 - t0).count();
        printf("threads %d: enqueue %.2f us/launch aggregate (%.0f k launches/s), incl. drain %.2f us/launch\n", T,
               1e6 * enq / (n * T), n * T / enq / 1e3, 1e6 * all / (n * T));
        for (auto& s : st) cudaStreamDestroy(s);
    }
    return 0;
}
