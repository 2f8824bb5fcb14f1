// common.cuh — shared host/device helpers for libaprb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/aprb200.h"

namespace aprb {

// ---- error plumbing (never throw across the C ABI) ------------------------------------------------------------
void set_error(const char* fmt, ...);

#define APRB_CUDA_OK(expr)                                                                          \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            aprb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));  \
            return APRB_ERR_CUDA;                                                                   \
        }                                                                                           \
    } while (0)

#define APRB_LAUNCH_OK() APRB_CUDA_OK(cudaGetLastError())

#define APRB_REQUIRE(cond, msg)                                  \
    do {                                                         \
        if (!(cond)) {                                           \
            aprb::set_error("%s: %s", __func__, msg);            \
            return APRB_ERR_INVALID;                             \
        }                                                        \
    } while (0)

// ---- workspace bump allocator (256-byte aligned carving of one caller-provided device buffer) ------------------
struct Carver {
    char* base;
    size_t off, cap;
    Carver(void* p, size_t bytes) : base((char*)p), off(0), cap(bytes) {}
    template <typename T>
    T* take(size_t n) {
        size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
        T* r = (T*)(base ? base + off : nullptr);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= cap; }
};
static inline size_t align256(size_t b) { return (b + 255) & ~size_t(255); }

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

int sm_count();
int current_device();   // cudaGetDevice, 0 on error

// ---- launch accounting + optional per-kernel CUDA-event timing (aprb_prof_*) --------------------------------------
// Every kernel launch (or CUB call, with its kernel count) goes through APRB_TIMED: it bumps the launch counter and,
// when profiling is enabled, brackets the launch with cudaEvents on the launching stream.
struct ProfScope {
    int slot;
    cudaStream_t st;
    ProfScope(const char* name, cudaStream_t st, int nlaunch = 1);
    ~ProfScope();
};
#define APRB_TIMED(name, st, n, stmt)            \
    do {                                         \
        aprb::ProfScope _ps(name, st, n);        \
        stmt;                                    \
    } while (0)

// ---- order-preserving float <-> int encoding for atomicMin/atomicMax on floats ---------------------------------
// Two fp32 values -> packed fp16x2 (lo in the low half), round to nearest even, SATURATING to +-65504 (one
// F2FP.SATFINITE.F16.F32.PACK_AB): every fp16 operand / activation store of the tensor path goes through this, so a
// feature outside the fp16 range clamps instead of turning into inf and poisoning the accumulators.
__device__ __forceinline__ unsigned pack_half2_sat(float lo, float hi) {
    unsigned r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// Index of the batch element containing stacked row i: largest b with off[b] <= i (off has B+1 entries).
__device__ __forceinline__ int find_cloud(const int* __restrict__ off, int B, int i) {
    int lo = 0, hi = B;  // invariant: off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// Single-block setup: exclusive prefixes lensA[B] -> offA[B+1] and (optional) lensB[B] -> offB[B+1], bbox[B*6] reset to
// (+max, -max) in ordered-int encoding, and `nzero` ints at `zero` cleared. One launch instead of four.
// Per-cloud bounding boxes: bbox[b*6 + {0,1,2}] = ordered-int min xyz, {3,4,5} = ordered-int max xyz.
__global__ void setup_kernel(const int* __restrict__ lensA, int* __restrict__ offA, const int* __restrict__ lensB,
                             int* __restrict__ offB, int B, int* __restrict__ bbox, int* __restrict__ zero, int nzero);
__global__ void bbox_kernel(const float* __restrict__ pts, int N, const int* __restrict__ off, int B,
                            int* __restrict__ bbox);

// CUB wrappers (temp storage carved from the workspace)
size_t scan_temp_bytes(int n);
size_t sort_temp_bytes(int n);
int exclusive_scan_i32(const int* d_in, int* d_out, int n, void* d_temp, size_t temp_bytes, cudaStream_t st);
int sort_pairs_i32(const uint64_t* k_in, uint64_t* k_out, const int* v_in, int* v_out, int n, void* d_temp,
                   size_t temp_bytes, cudaStream_t st);
int sort_pairs_i32(const uint32_t* k_in, uint32_t* k_out, const int* v_in, int* v_out, int n, void* d_temp,
                   size_t temp_bytes, cudaStream_t st);

}  // namespace aprb
