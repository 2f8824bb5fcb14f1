// pool_norm.cu — K4 (strided max-pool / nearest-upsample gathers) and K6 (InstanceNorm [+residual] [+LeakyReLU]).
//
// K4 replaces max_pool / closest_pool (/root/reference/Predator_APR/models/blocks.py:86-102, :71-83): the shadow
// index Ns selects an all-zero feature row, which takes part in the max exactly like the reference's appended row.
// K6 replaces BatchNormBlock.forward (blocks.py:459-468; nn.InstanceNorm1d over all N rows of the stacked pair,
// eps 1e-5, biased variance, no affine, no running stats) fused with the LeakyReLU(0.1) / residual add that follow it
// in UnaryBlock (:496-510), SimpleBlock (:592) and ResnetBottleneckBlock (:669-681).
#include "common.cuh"
#include <cuda_fp16.h>

namespace aprb {

// 4 channels stored in fp16 (8 bytes) <-> float4. Activations kept in fp16 are TF32-rounded values (10-bit mantissa):
// the narrowing is exact for magnitudes in fp16's normal range.
__device__ __forceinline__ float4 load_half4(const __half* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store_half4(__half* p, const float4 v) {
    uint2 u;
    u.x = pack_half2_sat(v.x, v.y); u.y = pack_half2_sat(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
}

// Columns of the index row that take part in the max: all H, or the reference's pool-matrix width min(max_count, limit)
// (neighbors.cpp:296-304 + dataloader.py:66-70) when the caller knows it — one scalar for a single collate, or one
// value per segment (collated pair) of a super-batch, looked up through the segments' query-row offsets.
__device__ __forceinline__ int pool_row_width(const int* __restrict__ d_width, const int* __restrict__ seg_off, int S, int n, int H) {
    if (!d_width) return H;
    int s = 0;
    if (seg_off) {
        while (s + 1 < S && n >= seg_off[s + 1]) ++s;
    }
    return min(H, max(d_width[s], 1));   // an all-pad segment keeps one (pad) column: the zero row
}

template <typename IdxT>
__global__ void max_pool_kernel(const float* __restrict__ x, const IdxT* __restrict__ idx, int ld, int Nq, int Ns,
                                int H, int C, const int* __restrict__ d_width, const int* __restrict__ seg_off, int S,
                                float* __restrict__ out) {
    // one thread per (query, 4-channel group); C % 4 == 0 path uses float4
    const int groups = C >> 2;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)Nq * groups) return;
    int n = (int)(t / groups), gch = (int)(t % groups);
    int Hn = pool_row_width(d_width, seg_off, S, n, H);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    const IdxT* row = idx + (size_t)n * ld;
    for (int h = 0; h < Hn; ++h) {
        long long s = (long long)row[h];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (s >= 0 && s < Ns) v = *reinterpret_cast<const float4*>(x + (size_t)s * C + 4 * gch);
        m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
    *reinterpret_cast<float4*>(out + (size_t)n * C + 4 * gch) = m;
}

// Warp per query, C = 128 * NV channels (lane owns NV float4): the index row is read once, coalesced, and only the valid
// neighbours are visited (ballot + shuffle broadcast) — pooling lists are ~half pad — with NV independent 128-bit loads
// in flight per neighbour. The thread-per-(query, quad) kernel above re-read the row once per quad and walked the pads.
template <typename IdxT, int NV, bool H16>
__global__ void __launch_bounds__(256)
max_pool_warp_kernel(const float* __restrict__ x, const IdxT* __restrict__ idx, int ld, int Nq, int Ns, int H,
                     const int* __restrict__ d_width, const int* __restrict__ seg_off, int S, float* __restrict__ out) {
    constexpr int C = 128 * NV;
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= Nq) return;
    const int Hn = pool_row_width(d_width, seg_off, S, n, H);
    float4 m[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) m[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    bool any_pad = false;
    const IdxT* row = idx + (size_t)n * ld;
    const float4* xb = reinterpret_cast<const float4*>(x) + lane;                       // fp32 rows: C/4 float4
    const __half* xh = reinterpret_cast<const __half*>(x) + lane * 4;                   // fp16 rows: C halves
    for (int h0 = 0; h0 < Hn; h0 += 32) {
        const int h = h0 + lane;
        long long sv = -1;
        if (h < Hn) sv = (long long)row[h];
        const bool valid = sv >= 0 && sv < Ns;
        unsigned vm = __ballot_sync(0xffffffffu, valid);
        any_pad |= vm != __ballot_sync(0xffffffffu, h < Hn);
        const int si = valid ? (int)sv : 0;
        while (vm) {
            const int src = __ffs(vm) - 1;
            vm &= vm - 1;
            const size_t srow = (size_t)__shfl_sync(0xffffffffu, si, src);
            const float4* p = xb + srow * (C / 4);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const float4 v = H16 ? load_half4(xh + srow * C + j * 128) : __ldg(p + j * 32);
                m[j].x = fmaxf(m[j].x, v.x); m[j].y = fmaxf(m[j].y, v.y); m[j].z = fmaxf(m[j].z, v.z); m[j].w = fmaxf(m[j].w, v.w);
            }
        }
    }
    float4* op = reinterpret_cast<float4*>(out + (size_t)n * C) + lane;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        if (any_pad) {                                               // the shadow neighbour's all-zero row takes part in the max
            m[j].x = fmaxf(m[j].x, 0.f); m[j].y = fmaxf(m[j].y, 0.f); m[j].z = fmaxf(m[j].z, 0.f); m[j].w = fmaxf(m[j].w, 0.f);
        }
        if (H16) store_half4(reinterpret_cast<__half*>(out) + (size_t)n * C + lane * 4 + j * 128, m[j]);
        else op[j * 32] = m[j];
    }
}

// fp16 rows of C = 256 * NV channels: lane owns NV groups of 8 channels (one 128-bit load each) and keeps the running
// maximum packed (HMNMX2: the max of fp16 values is exact, so nothing is widened) — 4 + NV * 5 instructions per valid
// neighbour against 4 + 2 NV * 9 for the float4 form above (the kernel was issue-bound, 923 warp instructions per query at
// C = 256, ncu round 2).
template <int NV>
__global__ void __launch_bounds__(256)
max_pool_h_kernel(const __half* __restrict__ x, const int* __restrict__ idx, int ld, int Nq, int Ns, int H,
                  const int* __restrict__ d_width, const int* __restrict__ seg_off, int S, __half* __restrict__ out) {
    constexpr int C = 256 * NV;
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= Nq) return;
    const int Hn = pool_row_width(d_width, seg_off, S, n, H);
    const __half2 ninf = __half2half2(__ushort_as_half((unsigned short)0xFC00u));
    __half2 m[NV][4];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) m[j][i] = ninf;
    bool any_pad = false;
    const int* row = idx + (size_t)n * ld;
    const uint4* xb = reinterpret_cast<const uint4*>(x) + lane;                         // row = C/8 uint4
    for (int h0 = 0; h0 < Hn; h0 += 32) {
        const int h = h0 + lane;
        int sv = -1;
        if (h < Hn) sv = row[h];
        const bool valid = sv >= 0 && sv < Ns;
        unsigned vm = __ballot_sync(0xffffffffu, valid);
        any_pad |= vm != __ballot_sync(0xffffffffu, h < Hn);
        const int si = valid ? sv : 0;
        while (vm) {
            const int src = __ffs(vm) - 1;
            vm &= vm - 1;
            const uint4* p = xb + (size_t)__shfl_sync(0xffffffffu, si, src) * (C / 8);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const uint4 v = __ldg(p + j * 32);
                m[j][0] = __hmax2(m[j][0], *reinterpret_cast<const __half2*>(&v.x));
                m[j][1] = __hmax2(m[j][1], *reinterpret_cast<const __half2*>(&v.y));
                m[j][2] = __hmax2(m[j][2], *reinterpret_cast<const __half2*>(&v.z));
                m[j][3] = __hmax2(m[j][3], *reinterpret_cast<const __half2*>(&v.w));
            }
        }
    }
    uint4* op = reinterpret_cast<uint4*>(out + (size_t)n * C) + lane;
    const __half2 zero = __half2half2(__ushort_as_half((unsigned short)0));
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        if (any_pad) {                                               // the shadow neighbour's all-zero row takes part in the max
#pragma unroll
            for (int i = 0; i < 4; ++i) m[j][i] = __hmax2(m[j][i], zero);
        }
        uint4 o;
        o.x = *reinterpret_cast<const unsigned*>(&m[j][0]); o.y = *reinterpret_cast<const unsigned*>(&m[j][1]);
        o.z = *reinterpret_cast<const unsigned*>(&m[j][2]); o.w = *reinterpret_cast<const unsigned*>(&m[j][3]);
        op[j * 32] = o;
    }
}

template <typename IdxT>
__global__ void max_pool_scalar_kernel(const float* __restrict__ x, const IdxT* __restrict__ idx, int ld, int Nq,
                                       int Ns, int H, int C, const int* __restrict__ d_width, const int* __restrict__ seg_off,
                                       int S, float* __restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)Nq * C) return;
    int n = (int)(t / C), c = (int)(t % C);
    int Hn = pool_row_width(d_width, seg_off, S, n, H);
    float m = -INFINITY;
    const IdxT* row = idx + (size_t)n * ld;
    for (int h = 0; h < Hn; ++h) {
        long long s = (long long)row[h];
        float v = (s >= 0 && s < Ns) ? x[(size_t)s * C + c] : 0.f;
        m = fmaxf(m, v);
    }
    out[(size_t)n * C + c] = m;
}

template <typename IdxT>
__global__ void closest_pool_kernel(const float* __restrict__ x, const IdxT* __restrict__ idx, int ld, int Nq, int Ns,
                                    int C, float* __restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)Nq * C) return;
    int n = (int)(t / C), c = (int)(t % C);
    long long s = (long long)idx[(size_t)n * ld];
    out[(size_t)n * C + c] = (s >= 0 && s < Ns) ? x[(size_t)s * C + c] : 0.f;
}

// Gradient of max_pool (training path): dy[n,c] goes to the FIRST neighbour that attains the maximum (torch.max's
// index on ties); the shadow row's share is dropped. One thread per (query, channel).
template <typename IdxT>
__global__ void max_pool_backward_kernel(const float* __restrict__ x, const IdxT* __restrict__ idx, int ld, int Nq, int Ns,
                                         int H, int C, const float* __restrict__ dy, float* __restrict__ dx) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)Nq * C) return;
    int n = (int)(t / C), c = (int)(t % C);
    const IdxT* row = idx + (size_t)n * ld;
    float m = -INFINITY;
    long long arg = Ns;
    for (int h = 0; h < H; ++h) {
        long long s = (long long)row[h];
        const bool real = s >= 0 && s < Ns;
        const float v = real ? x[(size_t)s * C + c] : 0.f;
        if (v > m) { m = v; arg = real ? s : Ns; }
    }
    if (arg < Ns) atomicAdd(dx + (size_t)arg * C + c, dy[(size_t)n * C + c]);
}

}  // namespace aprb

using namespace aprb;

extern "C" int aprb_max_pool_backward(const float* d_x, const void* d_idx, int idx_is_i64, int ld_idx, int Nq, int Ns,
                                      int H, int C, const float* d_dy, float* d_dx, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(Nq >= 0 && Ns >= 0 && H >= 1 && C >= 1 && ld_idx >= H, "bad shape");
    if (Ns == 0) return APRB_OK;
    APRB_REQUIRE(d_x && d_dx, "null pointer");
    APRB_CUDA_OK(cudaMemsetAsync(d_dx, 0, (size_t)Ns * C * sizeof(float), st));
    if (Nq == 0) return APRB_OK;
    APRB_REQUIRE(d_idx && d_dy, "null pointer");
    long long total = (long long)Nq * C;
    if (idx_is_i64) APRB_TIMED("max_pool_backward_kernel", st, 1, (max_pool_backward_kernel<long long><<<cdiv(total, 256), 256, 0, st>>>(d_x, (const long long*)d_idx, ld_idx, Nq, Ns, H, C, d_dy, d_dx)));
    else APRB_TIMED("max_pool_backward_kernel", st, 1, (max_pool_backward_kernel<int><<<cdiv(total, 256), 256, 0, st>>>(d_x, (const int*)d_idx, ld_idx, Nq, Ns, H, C, d_dy, d_dx)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

namespace aprb {

// ---- K6 -------------------------------------------------------------------------------------------------------
// Two launches per normalisation: (1) per (row-chunk, column) Welford partials (mean, M2); (2) every block of the apply
// kernel Chan-combines the partials of its 32 columns in a fixed order (deterministic, no float atomics), then
// standardises / adds the residual / applies LeakyReLU over its rows. The row-chunk size is picked on the host so that
// both grids fill the 148 SMs.
__device__ __forceinline__ void chan_combine(float& am, float& a2, int& an, float bm, float b2, int bn) {
    if (bn == 0) return;
    int n = an + bn;
    float d = bm - am;
    am += d * ((float)bn / (float)n);
    a2 += b2 + d * d * ((float)an * (float)bn / (float)n);
    an = n;
}

__global__ void __launch_bounds__(256)
norm_partial_kernel(const float* __restrict__ x, int N, int C, int rows_per_chunk, float* __restrict__ pmean,
                    float* __restrict__ pm2) {
    // grid: (ceil(C/32), chunks); block: (32, 8). Thread (tx, ty) walks rows r0+ty, r0+ty+8, ... of column c.
    __shared__ float s_mean[8][33], s_m2[8][33];
    __shared__ int s_cnt[8][33];
    int c = blockIdx.x * 32 + threadIdx.x;
    int r0 = blockIdx.y * rows_per_chunk, r1 = min(r0 + rows_per_chunk, N);
    float mean = 0.f, m2 = 0.f;
    int cnt = 0;
    if (c < C) {
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            float v = x[(size_t)r * C + c];
            ++cnt;
            float d = v - mean;
            mean += __fdividef(d, (float)cnt);
            m2 += d * (v - mean);
        }
    }
    s_mean[threadIdx.y][threadIdx.x] = mean; s_m2[threadIdx.y][threadIdx.x] = m2; s_cnt[threadIdx.y][threadIdx.x] = cnt;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float am = 0.f, a2 = 0.f;
        int an = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) chan_combine(am, a2, an, s_mean[k][threadIdx.x], s_m2[k][threadIdx.x], s_cnt[k][threadIdx.x]);
        pmean[(size_t)blockIdx.y * C + c] = am;
        pm2[(size_t)blockIdx.y * C + c] = a2;
    }
}

// combine `chunks` partials of one column; ty strides over chunks, then the 8 ty-partials are combined in order
__device__ __forceinline__ void block_column_stats(const float* __restrict__ pmean, const float* __restrict__ pm2, int N,
                                                   int C, int chunks, int rows_per_chunk, float eps, int c,
                                                   float (*s_a)[33], float (*s_b)[33], int (*s_n)[33],
                                                   float* s_mean, float* s_rstd) {
    float am = 0.f, a2 = 0.f;
    int an = 0;
    if (c < C)
        for (int k = threadIdx.y; k < chunks; k += 8)
            chan_combine(am, a2, an, pmean[(size_t)k * C + c], pm2[(size_t)k * C + c],
                         min(rows_per_chunk, N - k * rows_per_chunk));
    s_a[threadIdx.y][threadIdx.x] = am; s_b[threadIdx.y][threadIdx.x] = a2; s_n[threadIdx.y][threadIdx.x] = an;
    __syncthreads();
    if (threadIdx.y == 0) {
        float m = 0.f, v2 = 0.f;
        int n = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) chan_combine(m, v2, n, s_a[k][threadIdx.x], s_b[k][threadIdx.x], s_n[k][threadIdx.x]);
        s_mean[threadIdx.x] = m;
        s_rstd[threadIdx.x] = rsqrtf(v2 / (float)N + eps);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
norm_apply_kernel(const float* __restrict__ x, int N, int C, int chunks, int rows_per_chunk, float eps,
                  const float* __restrict__ pmean, const float* __restrict__ pm2, const float* __restrict__ res,
                  const float* __restrict__ rpmean, const float* __restrict__ rpm2, float slope, int rows_per_block,
                  int round_tf32, float* __restrict__ y) {
    // grid: (ceil(C/32), row blocks); block (32, 8)
    __shared__ float s_a[8][33], s_b[8][33];
    __shared__ int s_n[8][33];
    __shared__ float s_mean[32], s_rstd[32], s_rmean[32], s_rrstd[32];
    const int c = blockIdx.x * 32 + threadIdx.x;
    block_column_stats(pmean, pm2, N, C, chunks, rows_per_chunk, eps, c, s_a, s_b, s_n, s_mean, s_rstd);
    const bool nres = res != nullptr && rpmean != nullptr;
    if (nres) block_column_stats(rpmean, rpm2, N, C, chunks, rows_per_chunk, eps, c, s_a, s_b, s_n, s_rmean, s_rrstd);
    if (c >= C) return;
    const float mean = s_mean[threadIdx.x], rstd = s_rstd[threadIdx.x];
    const float rmean = nres ? s_rmean[threadIdx.x] : 0.f, rrstd = nres ? s_rrstd[threadIdx.x] : 1.f;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, N);
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
        const size_t i = (size_t)r * C + c;
        float v = (x[i] - mean) * rstd;
        if (res) v += (res[i] - rmean) * rrstd;
        v = v >= 0.f ? v : v * slope;
        if (round_tf32) {   // activations stored TF32-representable: downstream tensor-core GEMMs then read them exactly
            unsigned u;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
            v = __uint_as_float(u);
        }
        y[i] = v;
    }
}

}  // namespace aprb

using namespace aprb;

// last valid column + 1 of every index row, max-reduced per segment: the reference's pool-matrix width
template <typename IdxT>
__global__ void pool_seg_width_kernel(const IdxT* __restrict__ idx, int ld, int Nq, int Ns, int H, const int* __restrict__ seg_off,
                                      int S, int* __restrict__ seg_width) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= Nq) return;
    int last = 0;
    for (int h = lane; h < H; h += 32) {
        const long long v = (long long)idx[(size_t)n * ld + h];
        if (v >= 0 && v < Ns) last = h + 1;
    }
    last = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0 && last > 0) {
        int s = 0;
        if (seg_off) { while (s + 1 < S && n >= seg_off[s + 1]) ++s; }
        atomicMax(seg_width + s, last);
    }
}

extern "C" int aprb_pool_seg_widths(const void* d_idx, int idx_is_i64, int ld_idx, int Nq, int Ns, int H,
                                    const int32_t* d_seg_off, int S, int32_t* d_seg_width, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(Nq >= 0 && Ns >= 0 && H >= 1 && ld_idx >= H && S >= 1 && d_seg_width, "bad argument");
    APRB_REQUIRE(S == 1 || d_seg_off, "segment offsets missing");
    APRB_CUDA_OK(cudaMemsetAsync(d_seg_width, 0, sizeof(int) * (size_t)S, st));
    if (Nq == 0) return APRB_OK;
    APRB_REQUIRE(d_idx, "null pointer");
    if (idx_is_i64) APRB_TIMED("pool_seg_width_kernel", st, 1, (pool_seg_width_kernel<long long><<<cdiv(Nq, 8), 256, 0, st>>>((const long long*)d_idx, ld_idx, Nq, Ns, H, d_seg_off, S, d_seg_width)));
    else APRB_TIMED("pool_seg_width_kernel", st, 1, (pool_seg_width_kernel<int><<<cdiv(Nq, 8), 256, 0, st>>>((const int*)d_idx, ld_idx, Nq, Ns, H, d_seg_off, S, d_seg_width)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

static int max_pool_impl(const void* d_xv, int x_is_f16, const void* d_idx, int idx_is_i64, int ld_idx, int Nq, int Ns, int H,
                         int C, const int32_t* d_width, const int32_t* d_seg_off, int S, void* d_outv, cudaStream_t st) {
    APRB_REQUIRE(Nq >= 0 && Ns >= 0 && H >= 1 && C >= 1 && ld_idx >= H, "bad shape");
    APRB_REQUIRE(!d_seg_off || (d_width && S >= 1), "segment offsets need per-segment widths");
    if (Nq == 0) return APRB_OK;
    APRB_REQUIRE(d_idx && d_outv && (d_xv || Ns == 0), "null pointer");
    const float* d_x = (const float*)d_xv;
    float* d_out = (float*)d_outv;
    const int T = 256;
    if (x_is_f16) {
        APRB_REQUIRE(C >= 128 && C % 128 == 0 && C <= 1024 && !idx_is_i64, "fp16 max_pool needs C in {128, ..., 1024} and int32 indices");
        if (C % 256 == 0 && (((uintptr_t)d_x | (uintptr_t)d_out) & 15) == 0) {
#define MPH2(NV) APRB_TIMED("max_pool_kernel", st, 1, (max_pool_h_kernel<NV><<<cdiv(Nq, 8), 256, 0, st>>>((const __half*)d_x, (const int*)d_idx, ld_idx, Nq, Ns, H, d_width, d_seg_off, S, (__half*)d_out)))
            switch (C / 256) { case 1: MPH2(1); break; case 2: MPH2(2); break; case 3: MPH2(3); break; default: MPH2(4); break; }
#undef MPH2
            APRB_LAUNCH_OK();
            return APRB_OK;
        }
#define MPH(NV) APRB_TIMED("max_pool_kernel", st, 1, (max_pool_warp_kernel<int, NV, true><<<cdiv(Nq, 8), 256, 0, st>>>(d_x, (const int*)d_idx, ld_idx, Nq, Ns, H, d_width, d_seg_off, S, d_out)))
        switch (C / 128) {
            case 1: MPH(1); break; case 2: MPH(2); break; case 3: MPH(3); break; case 4: MPH(4); break;
            case 5: MPH(5); break; case 6: MPH(6); break; case 7: MPH(7); break; default: MPH(8); break;
        }
#undef MPH
        APRB_LAUNCH_OK();
        return APRB_OK;
    }
    if (C % 4 == 0 && ((uintptr_t)d_x % 16 == 0) && ((uintptr_t)d_out % 16 == 0)) {
        long long total = (long long)Nq * (C / 4);
        if (C % 128 == 0 && C <= 1024 && H >= 1) {
#define MPW(IDX, NV) APRB_TIMED("max_pool_kernel", st, 1, (max_pool_warp_kernel<IDX, NV, false><<<cdiv(Nq, 8), 256, 0, st>>>(d_x, (const IDX*)d_idx, ld_idx, Nq, Ns, H, d_width, d_seg_off, S, d_out)))
#define MPW_NV(IDX) do { if (C == 128) MPW(IDX, 1); else if (C == 256) MPW(IDX, 2); else if (C == 384) MPW(IDX, 3); else if (C == 512) MPW(IDX, 4); \
                         else if (C == 640) MPW(IDX, 5); else if (C == 768) MPW(IDX, 6); else if (C == 896) MPW(IDX, 7); else MPW(IDX, 8); } while (0)
            if (idx_is_i64) MPW_NV(long long); else MPW_NV(int);
#undef MPW_NV
#undef MPW
            APRB_LAUNCH_OK();
            return APRB_OK;
        }
        if (idx_is_i64) APRB_TIMED("max_pool_kernel", st, 1, (max_pool_kernel<long long><<<cdiv(total, T), T, 0, st>>>(d_x, (const long long*)d_idx, ld_idx, Nq, Ns, H, C, d_width, d_seg_off, S, d_out)));
        else APRB_TIMED("max_pool_kernel", st, 1, (max_pool_kernel<int><<<cdiv(total, T), T, 0, st>>>(d_x, (const int*)d_idx, ld_idx, Nq, Ns, H, C, d_width, d_seg_off, S, d_out)));
    } else {
        long long total = (long long)Nq * C;
        if (idx_is_i64) APRB_TIMED("max_pool_scalar_kernel", st, 1, (max_pool_scalar_kernel<long long><<<cdiv(total, T), T, 0, st>>>(d_x, (const long long*)d_idx, ld_idx, Nq, Ns, H, C, d_width, d_seg_off, S, d_out)));
        else APRB_TIMED("max_pool_scalar_kernel", st, 1, (max_pool_scalar_kernel<int><<<cdiv(total, T), T, 0, st>>>(d_x, (const int*)d_idx, ld_idx, Nq, Ns, H, C, d_width, d_seg_off, S, d_out)));
    }
    APRB_LAUNCH_OK();
    return APRB_OK;
}

extern "C" int aprb_max_pool(const float* d_x, const void* d_idx, int idx_is_i64, int ld_idx, int Nq, int Ns, int H,
                             int C, const int32_t* d_width, float* d_out, void* stream) {
    return max_pool_impl(d_x, 0, d_idx, idx_is_i64, ld_idx, Nq, Ns, H, C, d_width, nullptr, 1, d_out, (cudaStream_t)stream);
}

extern "C" int aprb_max_pool_f16(const void* d_x16, const int32_t* d_idx, int ld_idx, int Nq, int Ns, int H, int C,
                                 void* d_out16, void* stream) {
    return max_pool_impl(d_x16, 1, d_idx, 0, ld_idx, Nq, Ns, H, C, nullptr, nullptr, 1, d_out16, (cudaStream_t)stream);
}

extern "C" int aprb_max_pool_seg(const void* d_x, int x_is_f16, const void* d_idx, int idx_is_i64, int ld_idx, int Nq, int Ns,
                                 int H, int C, const int32_t* d_seg_off, int S, const int32_t* d_seg_width, void* d_out,
                                 void* stream) {
    return max_pool_impl(d_x, x_is_f16, d_idx, idx_is_i64, ld_idx, Nq, Ns, H, C, d_seg_width, S > 1 ? d_seg_off : nullptr, S,
                         d_out, (cudaStream_t)stream);
}

__global__ void f32_to_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const unsigned u = pack_half2_sat(in[i], 0.f); out[i] = *reinterpret_cast<const __half*>(&u); }
}

extern "C" int aprb_f32_to_f16(const float* d_in, void* d_out16, size_t n, void* stream) {
    APRB_REQUIRE(n == 0 || (d_in && d_out16), "null pointer");
    if (n == 0) return APRB_OK;
    APRB_TIMED("f32_to_f16_kernel", (cudaStream_t)stream, 1, (f32_to_f16_kernel<<<cdiv((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(d_in, (__half*)d_out16, n)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

extern "C" int aprb_closest_pool(const float* d_x, const void* d_idx, int idx_is_i64, int ld_idx, int Nq, int Ns, int C,
                                 float* d_out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(Nq >= 0 && Ns >= 0 && C >= 1 && ld_idx >= 1, "bad shape");
    if (Nq == 0) return APRB_OK;
    APRB_REQUIRE(d_idx && d_out && (d_x || Ns == 0), "null pointer");
    const int T = 256;
    long long total = (long long)Nq * C;
    if (idx_is_i64) APRB_TIMED("closest_pool_kernel", st, 1, (closest_pool_kernel<long long><<<cdiv(total, T), T, 0, st>>>(d_x, (const long long*)d_idx, ld_idx, Nq, Ns, C, d_out)));
    else APRB_TIMED("closest_pool_kernel", st, 1, (closest_pool_kernel<int><<<cdiv(total, T), T, 0, st>>>(d_x, (const int*)d_idx, ld_idx, Nq, Ns, C, d_out)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

// ---- float4 variants (C % 4 == 0, C/4 a power of two <= 256 or a multiple of 256): 16-byte accesses, shifted
// sums instead of a serial Welford chain, so each thread keeps several independent loads in flight.
__global__ void __launch_bounds__(256)
norm_partial4_kernel(const float* __restrict__ x, const float* __restrict__ x2, int N, int C, int W, int rows_per_chunk,
                     size_t tensor_stride, float* __restrict__ pmean, float* __restrict__ pm2) {
    if (blockIdx.z == 1) { x = x2; pmean += tensor_stride; pm2 += tensor_stride; }   // second tensor (normalised residual)
    __shared__ float4 s_s1[256], s_s2[256];
    __shared__ int s_cnt[256];
    const int Cq = C >> 2, R = 256 / W;
    const int tx = threadIdx.x % W, ty = threadIdx.x / W;
    const int quad = blockIdx.x * W + tx;
    const int r0 = blockIdx.y * rows_per_chunk, r1 = min(r0 + rows_per_chunk, N);
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1, pv = s1;
    int cnt = 0;
    if (quad < Cq) {
        const float4* xp = reinterpret_cast<const float4*>(x) + quad;
        pv = xp[(size_t)r0 * Cq];                                  // pivot: first row of the chunk (shifted-data sums)
#pragma unroll 4
        for (int r = r0 + ty; r < r1; r += R) {
            const float4 v = xp[(size_t)r * Cq];
            const float a = v.x - pv.x, b = v.y - pv.y, c = v.z - pv.z, d = v.w - pv.w;
            s1.x += a; s1.y += b; s1.z += c; s1.w += d;
            s2.x = fmaf(a, a, s2.x); s2.y = fmaf(b, b, s2.y); s2.z = fmaf(c, c, s2.z); s2.w = fmaf(d, d, s2.w);
            ++cnt;
        }
    }
    s_s1[threadIdx.x] = s1; s_s2[threadIdx.x] = s2; s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    if (ty == 0 && quad < Cq) {
        float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
        int n = 0;
        for (int k = 0; k < R; ++k) {                              // fixed order -> deterministic
            const float4 b1 = s_s1[k * W + tx], b2 = s_s2[k * W + tx];
            a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
            a2.x += b2.x; a2.y += b2.y; a2.z += b2.z; a2.w += b2.w;
            n += s_cnt[k * W + tx];
        }
        const float inv = 1.0f / (float)max(n, 1);
        float4 mean = make_float4(pv.x + a1.x * inv, pv.y + a1.y * inv, pv.z + a1.z * inv, pv.w + a1.w * inv);
        float4 m2 = make_float4(fmaxf(a2.x - a1.x * a1.x * inv, 0.f), fmaxf(a2.y - a1.y * a1.y * inv, 0.f),
                                fmaxf(a2.z - a1.z * a1.z * inv, 0.f), fmaxf(a2.w - a1.w * a1.w * inv, 0.f));
        reinterpret_cast<float4*>(pmean + (size_t)blockIdx.y * C)[quad] = mean;
        reinterpret_cast<float4*>(pm2 + (size_t)blockIdx.y * C)[quad] = m2;
    }
}

// One warp per column: lanes stride over the chunk partials, then a fixed-shape shuffle tree (deterministic).
// grid (ceil(C/8), ntensors); block 256. stats layout: [tensor][2][C] = mean, rstd.
__global__ void __launch_bounds__(256)
norm_finalize_kernel2(const float* __restrict__ pmean, const float* __restrict__ pm2, int N, int C, int chunks,
                      int rows_per_chunk, float eps, size_t tensor_stride, float* __restrict__ stats) {
    const int col = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (col >= C) return;
    pmean += blockIdx.y * tensor_stride; pm2 += blockIdx.y * tensor_stride;
    float am = 0.f, a2 = 0.f;
    int an = 0;
    for (int k = lane; k < chunks; k += 32)
        chan_combine(am, a2, an, pmean[(size_t)k * C + col], pm2[(size_t)k * C + col], min(rows_per_chunk, N - k * rows_per_chunk));
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float bm = __shfl_xor_sync(0xffffffffu, am, d), b2 = __shfl_xor_sync(0xffffffffu, a2, d);
        const int bn = __shfl_xor_sync(0xffffffffu, an, d);
        // combine (lower lane, upper lane) in that order on both sides so every lane ends with the same value
        float lm = (lane & d) ? bm : am, l2 = (lane & d) ? b2 : a2; int ln = (lane & d) ? bn : an;
        const float um = (lane & d) ? am : bm, u2 = (lane & d) ? a2 : b2; const int un = (lane & d) ? an : bn;
        chan_combine(lm, l2, ln, um, u2, un);
        am = lm; a2 = l2; an = ln;
    }
    if (lane == 0) {
        stats[(size_t)blockIdx.y * 2 * C + col] = am;
        stats[(size_t)blockIdx.y * 2 * C + C + col] = rsqrtf(a2 / (float)N + eps);
    }
}

__device__ __forceinline__ float act_round(float v, float slope, int round_tf32) {
    v = v >= 0.f ? v : v * slope;
    if (round_tf32) {
        unsigned u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
        v = __uint_as_float(u);
    }
    return v;
}

__global__ void __launch_bounds__(256)
norm_apply4_kernel(const float* __restrict__ x, int N, int C, int W, const float* __restrict__ stats,
                   const float* __restrict__ res, int norm_res, float slope, int rows_per_block, int round_tf32,
                   float* __restrict__ y) {
    const int Cq = C >> 2, R = 256 / W;
    const int tx = threadIdx.x % W, ty = threadIdx.x / W;
    const int quad = blockIdx.x * W + tx;
    if (quad >= Cq) return;
    const float4 mean = reinterpret_cast<const float4*>(stats)[quad];
    const float4 rstd = reinterpret_cast<const float4*>(stats + C)[quad];
    float4 rmean = make_float4(0.f, 0.f, 0.f, 0.f), rrstd = make_float4(1.f, 1.f, 1.f, 1.f);
    if (res && norm_res) {
        rmean = reinterpret_cast<const float4*>(stats + 2 * (size_t)C)[quad];
        rrstd = reinterpret_cast<const float4*>(stats + 3 * (size_t)C)[quad];
    }
    const int r0 = blockIdx.y * rows_per_block, r1 = min(r0 + rows_per_block, N);
    const float4* xp = reinterpret_cast<const float4*>(x) + quad;
    const float4* rp = res ? reinterpret_cast<const float4*>(res) + quad : nullptr;
    float4* yp = reinterpret_cast<float4*>(y) + quad;
#pragma unroll 4
    for (int r = r0 + ty; r < r1; r += R) {
        const float4 v = xp[(size_t)r * Cq];
        float4 o = make_float4((v.x - mean.x) * rstd.x, (v.y - mean.y) * rstd.y, (v.z - mean.z) * rstd.z, (v.w - mean.w) * rstd.w);
        if (rp) {
            const float4 q = rp[(size_t)r * Cq];
            o.x += (q.x - rmean.x) * rrstd.x; o.y += (q.y - rmean.y) * rrstd.y;
            o.z += (q.z - rmean.z) * rrstd.z; o.w += (q.w - rmean.w) * rrstd.w;
        }
        o.x = act_round(o.x, slope, round_tf32); o.y = act_round(o.y, slope, round_tf32);
        o.z = act_round(o.z, slope, round_tf32); o.w = act_round(o.w, slope, round_tf32);
        yp[(size_t)r * Cq] = o;
    }
}

static void norm_plan(int N, int C, int* rows_per_chunk, int* chunks) {
    // about 2 waves of (column-tile x chunk) blocks, chunks of at least 64 rows, at most 256 chunks
    int ct = cdiv(C, 32);
    int want = max(1, (2 * sm_count() + ct - 1) / ct);
    int ch = min(min(want, 256), max(1, N / 64));
    int rpc = ((cdiv(N, ch) + 7) / 8) * 8;
    *rows_per_chunk = rpc;
    *chunks = cdiv(N, rpc);
}

extern "C" size_t aprb_instnorm_ws_bytes(int N, int C) {
    if (N < 0 || C < 0) return 0;
    return 4 * align256((size_t)256 * (C > 0 ? C : 1) * sizeof(float)) + align256((size_t)4 * (C > 0 ? C : 1) * sizeof(float)) + 512;
}

extern "C" int aprb_instnorm_lrelu(const float* d_x, int N, int C, float eps, float slope, const float* d_residual,
                                   int norm_residual, int round_tf32, float* d_y, void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(N >= 0 && C >= 1, "bad shape");
    if (N == 0) return APRB_OK;
    APRB_REQUIRE(d_x && d_y && d_ws, "null pointer");
    if (ws_bytes < aprb_instnorm_ws_bytes(N, C)) { set_error("aprb_instnorm_lrelu: workspace too small"); return APRB_ERR_WORKSPACE; }
    Carver c(d_ws, ws_bytes);
    // one contiguous block: pmean | rpmean | pm2 | rpm2 | stats  (each 256*C floats; the residual's partials sit one
    // "tensor stride" (= 2 arrays) after the main tensor's in both the mean and the M2 halves)
    float* blockp = c.take<float>((size_t)4 * 256 * C + (size_t)4 * C);
    float* pmean = blockp; float* pm2 = blockp + (size_t)256 * C;
    float* rpmean = blockp + (size_t)2 * 256 * C; float* rpm2 = blockp + (size_t)3 * 256 * C;
    int rpc, chunks;
    const bool nr = d_residual && norm_residual;
    const int Cq = C / 4;
    const bool aligned = (((uintptr_t)d_x | (uintptr_t)d_y | (uintptr_t)(d_residual ? d_residual : d_x)) & 15) == 0;
    if (C % 4 == 0 && aligned && ((Cq <= 256 && (Cq & (Cq - 1)) == 0) || Cq % 256 == 0)) {
        const int W = Cq < 256 ? Cq : 256, gx = cdiv(Cq, W);
        const int sms = sm_count();
        // statistics pass: ~4 waves of blocks, chunks of >= 32 rows, <= 256 chunks
        int ch = min(min(max(1, 4 * sms / (gx * (nr ? 2 : 1))), 256), max(1, N / 32));
        rpc = cdiv(N, ch);
        chunks = cdiv(N, rpc);
        const size_t tstride = (size_t)256 * C;   // pmean/pm2 of the residual live one tensor-stride further (= rpmean/rpm2)
        float* stats = rpm2 + tstride;            // [2 tensors][mean, rstd][C] (carved below)
        APRB_TIMED("norm_partial4_kernel", st, 1, (norm_partial4_kernel<<<dim3(gx, chunks, nr ? 2 : 1), 256, 0, st>>>(
            d_x, d_residual, N, C, W, rpc, 2 * tstride, pmean, pm2)));
        APRB_TIMED("norm_finalize_kernel2", st, 1, (norm_finalize_kernel2<<<dim3(cdiv(C, 8), nr ? 2 : 1), 256, 0, st>>>(
            pmean, pm2, N, C, chunks, rpc, eps, 2 * tstride, stats)));
        // apply pass: its own row tiling, ~6 waves of blocks
        int rb = min(max(1, 6 * sms / gx), max(1, N / 16));
        int rpb = cdiv(N, rb);
        APRB_TIMED("norm_apply4_kernel", st, 1, (norm_apply4_kernel<<<dim3(gx, cdiv(N, rpb)), 256, 0, st>>>(
            d_x, N, C, W, stats, d_residual, nr ? 1 : 0, slope, rpb, round_tf32, d_y)));
        APRB_LAUNCH_OK();
        return APRB_OK;
    }
    norm_plan(N, C, &rpc, &chunks);
    const dim3 blk(32, 8);
    APRB_TIMED("norm_partial_kernel", st, 1, (norm_partial_kernel<<<dim3(cdiv(C, 32), chunks), blk, 0, st>>>(d_x, N, C, rpc, pmean, pm2)));
    if (nr) APRB_TIMED("norm_partial_kernel", st, 1, (norm_partial_kernel<<<dim3(cdiv(C, 32), chunks), blk, 0, st>>>(d_residual, N, C, rpc, rpmean, rpm2)));
    int rpb = rpc;   // same row tiling for the apply pass
    APRB_TIMED("norm_apply_kernel", st, 1, (norm_apply_kernel<<<dim3(cdiv(C, 32), cdiv(N, rpb)), blk, 0, st>>>(
        d_x, N, C, chunks, rpc, eps, pmean, pm2, d_residual, nr ? rpmean : nullptr, nr ? rpm2 : nullptr, slope, rpb, round_tf32, d_y)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

// ---- segmented variant (super-batched pairs) ---------------------------------------------------------------------
// P stacked pairs share every other kernel of the path, but BatchNormBlock (= InstanceNorm over all rows of ONE
// collated pair, blocks.py:459-468) must keep per-pair statistics: rows [seg_off[s], seg_off[s+1]) form segment s.
// Two launches: (1) per (segment, chunk, column-quad) shifted-sum partials; (2) every apply block Chan-combines the
// `ch` partials of its columns in a fixed order (deterministic), then standardises its rows. Segment bounds are read
// on the device (the host never learns the per-level cloud lengths), chunking is relative to each segment.
namespace aprb {

__device__ __forceinline__ void chan_combine4(float4& am, float4& a2, int& an, const float4 bm, const float4 b2, int bn) {
    if (bn <= 0) return;
    const int n = an + bn;
    const float fb = (float)bn / (float)n, fab = (float)an * (float)bn / (float)n;
    float d;
    d = bm.x - am.x; am.x += d * fb; a2.x += b2.x + d * d * fab;
    d = bm.y - am.y; am.y += d * fb; a2.y += b2.y + d * d * fab;
    d = bm.z - am.z; am.z += d * fb; a2.z += b2.z + d * d * fab;
    d = bm.w - am.w; am.w += d * fb; a2.w += b2.w + d * d * fab;
    an = n;
}

// grid (gx, ch, S * nt); block 256 = W quads x R rows. The LAST block to finish a (segment, tensor, column tile) —
// a ticket counter decides who is last, never the order of the sum — Chan-combines the ch partials in index order and
// writes stats[(seg*nt + t)][mean | rstd][C].
__global__ void __launch_bounds__(256)
norm_seg_partial_kernel(const float* __restrict__ x, const float* __restrict__ x2, const int* __restrict__ seg_off, int N,
                        int C, int W, int ch, int nt, int t_first, int t_count, float eps, float* __restrict__ pmean,
                        float* __restrict__ pm2, int* __restrict__ tickets, float* __restrict__ stats) {
    // blockIdx.z = seg * t_count + tt covers tensors t_first .. t_first + t_count - 1 of the nt normalised tensors
    const int seg = blockIdx.z / t_count, t = t_first + (blockIdx.z - seg * t_count);
    if (t == 1) x = x2;
    __shared__ float4 s_s1[256], s_s2[256];
    __shared__ int s_cnt[256];
    __shared__ int s_last;
    const int Cq = C >> 2, R = 256 / W;
    const int tx = threadIdx.x % W, ty = threadIdx.x / W;
    const int quad = blockIdx.x * W + tx;
    const int a = seg_off ? seg_off[seg] : 0, b = seg_off ? seg_off[seg + 1] : N;
    const int rpc = (b - a + ch - 1) / ch;
    const int r0 = a + blockIdx.y * rpc, r1 = min(r0 + rpc, b);
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1, pv = s1;
    int cnt = 0;
    if (quad < Cq && r0 < r1) {
        const float4* xp = reinterpret_cast<const float4*>(x) + quad;
        pv = xp[(size_t)r0 * Cq];                                  // pivot: first row of the chunk (shifted-data sums)
#pragma unroll 4
        for (int r = r0 + ty; r < r1; r += R) {
            const float4 v = xp[(size_t)r * Cq];
            const float e = v.x - pv.x, f = v.y - pv.y, g = v.z - pv.z, h = v.w - pv.w;
            s1.x += e; s1.y += f; s1.z += g; s1.w += h;
            s2.x = fmaf(e, e, s2.x); s2.y = fmaf(f, f, s2.y); s2.z = fmaf(g, g, s2.z); s2.w = fmaf(h, h, s2.w);
            ++cnt;
        }
    }
    s_s1[threadIdx.x] = s1; s_s2[threadIdx.x] = s2; s_cnt[threadIdx.x] = cnt;
    __syncthreads();
    const size_t slot0 = (size_t)blockIdx.z * ch * C;
    if (ty == 0 && quad < Cq) {
        float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;
        int n = 0;
        for (int k = 0; k < R; ++k) {                              // fixed order -> deterministic
            const float4 b1 = s_s1[k * W + tx], b2 = s_s2[k * W + tx];
            a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
            a2.x += b2.x; a2.y += b2.y; a2.z += b2.z; a2.w += b2.w;
            n += s_cnt[k * W + tx];
        }
        const float inv = 1.0f / (float)max(n, 1);
        const float4 mean = make_float4(pv.x + a1.x * inv, pv.y + a1.y * inv, pv.z + a1.z * inv, pv.w + a1.w * inv);
        const float4 m2 = make_float4(fmaxf(a2.x - a1.x * a1.x * inv, 0.f), fmaxf(a2.y - a1.y * a1.y * inv, 0.f),
                                      fmaxf(a2.z - a1.z * a1.z * inv, 0.f), fmaxf(a2.w - a1.w * a1.w * inv, 0.f));
        reinterpret_cast<float4*>(pmean + slot0 + (size_t)blockIdx.y * C)[quad] = mean;
        reinterpret_cast<float4*>(pm2 + slot0 + (size_t)blockIdx.y * C)[quad] = m2;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        int* tk = tickets + blockIdx.z * gridDim.x + blockIdx.x;
        const int old = atomicAdd(tk, 1);
        s_last = (old == ch - 1);
        if (s_last) *tk = 0;                                       // self-reset for the next use of this workspace
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // ---- finalize: chunk partials combined in chunk order (ty strides, then the R ty-partials in order)
    float4 am = make_float4(0.f, 0.f, 0.f, 0.f), a2 = am;
    int an = 0;
    if (quad < Cq)
        for (int k = ty; k < ch; k += R) {
            const int c2 = min(rpc, b - a - k * rpc);
            if (c2 <= 0) continue;
            chan_combine4(am, a2, an, __ldcg(reinterpret_cast<const float4*>(pmean + slot0 + (size_t)k * C) + quad),
                          __ldcg(reinterpret_cast<const float4*>(pm2 + slot0 + (size_t)k * C) + quad), c2);
        }
    s_s1[threadIdx.x] = am; s_s2[threadIdx.x] = a2; s_cnt[threadIdx.x] = an;
    __syncthreads();
    if (ty == 0 && quad < Cq) {
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f), v2 = m;
        int n = 0;
        for (int k = 0; k < R; ++k) chan_combine4(m, v2, n, s_s1[k * W + tx], s_s2[k * W + tx], s_cnt[k * W + tx]);
        const float inv = 1.0f / (float)max(b - a, 1);
        float* st = stats + (size_t)(seg * nt + t) * 2 * C;
        reinterpret_cast<float4*>(st)[quad] = m;
        reinterpret_cast<float4*>(st + C)[quad] = make_float4(rsqrtf(v2.x * inv + eps), rsqrtf(v2.y * inv + eps),
                                                              rsqrtf(v2.z * inv + eps), rsqrtf(v2.w * inv + eps));
    }
}

// grid (gx, rb, S); block 256 = W quads x R rows
__global__ void __launch_bounds__(256)
norm_seg_apply_kernel(const float* __restrict__ x, const int* __restrict__ seg_off, int N, int C, int W, int nt,
                      const float* __restrict__ stats, const float* __restrict__ res, float slope, int round_tf32,
                      float* __restrict__ y, int res16, int out16) {
    const int seg = blockIdx.z;
    const int Cq = C >> 2, R = 256 / W;
    const int tx = threadIdx.x % W, ty = threadIdx.x / W;
    const int quad = blockIdx.x * W + tx;
    const int a = seg_off ? seg_off[seg] : 0, b = seg_off ? seg_off[seg + 1] : N;
    if (a >= b || quad >= Cq) return;
    const float* st = stats + (size_t)seg * nt * 2 * C;
    const float4 mean = reinterpret_cast<const float4*>(st)[quad], rstd = reinterpret_cast<const float4*>(st + C)[quad];
    float4 rmean = make_float4(0.f, 0.f, 0.f, 0.f), rrstd = make_float4(1.f, 1.f, 1.f, 1.f);
    if (nt == 2) { rmean = reinterpret_cast<const float4*>(st + 2 * (size_t)C)[quad]; rrstd = reinterpret_cast<const float4*>(st + 3 * (size_t)C)[quad]; }
    const int rpb = (b - a + gridDim.y - 1) / gridDim.y;
    const int r0 = a + blockIdx.y * rpb, r1 = min(r0 + rpb, b);
    const float4* xp = reinterpret_cast<const float4*>(x) + quad;
    const float4* rp = res ? reinterpret_cast<const float4*>(res) + quad : nullptr;
    float4* yp = reinterpret_cast<float4*>(y) + quad;
#pragma unroll 4
    for (int r = r0 + ty; r < r1; r += R) {
        const float4 v = xp[(size_t)r * Cq];
        float4 o = make_float4((v.x - mean.x) * rstd.x, (v.y - mean.y) * rstd.y, (v.z - mean.z) * rstd.z, (v.w - mean.w) * rstd.w);
        if (rp) {
            const float4 q = res16 ? load_half4(reinterpret_cast<const __half*>(res) + ((size_t)r * Cq + quad) * 4) : rp[(size_t)r * Cq];
            o.x += (q.x - rmean.x) * rrstd.x; o.y += (q.y - rmean.y) * rrstd.y;
            o.z += (q.z - rmean.z) * rrstd.z; o.w += (q.w - rmean.w) * rrstd.w;
        }
        o.x = act_round(o.x, slope, round_tf32); o.y = act_round(o.y, slope, round_tf32);
        o.z = act_round(o.z, slope, round_tf32); o.w = act_round(o.w, slope, round_tf32);
        if (out16) store_half4(reinterpret_cast<__half*>(y) + ((size_t)r * Cq + quad) * 4, o);
        else yp[(size_t)r * Cq] = o;
    }
}

// Statistics from the producer's group partials (GEMM epilogue: per 32-row group mean and M2, see gemm_tcgen05.cu)
// instead of a pass over the tensor: for segment rows [a, b) the whole groups [ceil(a/32), floor(b/32)) are combined
// from `gs` (1/16 of the tensor's bytes) and the ragged head / tail rows (< 32 each) are read from the tensor itself, so
// any segmentation is handled exactly. grid (gx, S * t_count); block 256 = W quads x R lanes over groups; fixed order.
__global__ void __launch_bounds__(256)
norm_seg_groups_kernel(const float* __restrict__ x, const float* __restrict__ x2, const float* __restrict__ gs,
                       const float* __restrict__ gs2, const int* __restrict__ seg_off, int N, int C, int W, int nt,
                       int t_first, int t_count, float eps, float* __restrict__ stats) {
    const int seg = blockIdx.y / t_count, t = t_first + (blockIdx.y - seg * t_count);
    if (t == 1) { x = x2; gs = gs2; }
    __shared__ float4 s_m[256], s_v[256];
    __shared__ int s_n[256];
    const int Cq = C >> 2, R = 256 / W;
    const int tx = threadIdx.x % W, ty = threadIdx.x / W;
    const int quad = blockIdx.x * W + tx;
    const int a = seg_off ? seg_off[seg] : 0, b = seg_off ? seg_off[seg + 1] : N;
    int g0 = (a + 31) >> 5, g1 = b >> 5;
    int head_end = g0 << 5, tail_beg = g1 << 5;
    if (g0 >= g1) { g1 = g0; head_end = b; tail_beg = b; }           // no whole group inside the segment: rows only
    float4 am = make_float4(0.f, 0.f, 0.f, 0.f), a2 = am;
    int an = 0;
    if (quad < Cq) {
        const float4* xp = reinterpret_cast<const float4*>(x) + quad;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = a + ty; r < head_end; r += R) chan_combine4(am, a2, an, xp[(size_t)r * Cq], zero, 1);
        // whole groups all count 32 rows: shifted sums of the group means around the first one this thread sees
        // (mean = p + S1/G, M2 = sum M2_g + 32 (S2 - S1^2/G)), one Chan combine at the end instead of one per group
        const float4* gp = reinterpret_cast<const float4*>(gs) + quad;
        if (g0 + ty < g1) {
            const float4 pv = __ldcg(gp + (size_t)(g0 + ty) * 2 * Cq);
            float4 s1 = zero, s2 = zero, sm2 = zero;
            int G = 0;
#pragma unroll 8
            for (int g = g0 + ty; g < g1; g += R) {
                const float4 m = __ldcg(gp + (size_t)g * 2 * Cq), v = __ldcg(gp + ((size_t)g * 2 + 1) * Cq);
                const float e = m.x - pv.x, f = m.y - pv.y, gg = m.z - pv.z, hh = m.w - pv.w;
                s1.x += e; s1.y += f; s1.z += gg; s1.w += hh;
                s2.x = fmaf(e, e, s2.x); s2.y = fmaf(f, f, s2.y); s2.z = fmaf(gg, gg, s2.z); s2.w = fmaf(hh, hh, s2.w);
                sm2.x += v.x; sm2.y += v.y; sm2.z += v.z; sm2.w += v.w;
                ++G;
            }
            const float invG = 1.0f / (float)G;
            const float4 gm = make_float4(pv.x + s1.x * invG, pv.y + s1.y * invG, pv.z + s1.z * invG, pv.w + s1.w * invG);
            const float4 g2 = make_float4(sm2.x + 32.f * fmaxf(s2.x - s1.x * s1.x * invG, 0.f), sm2.y + 32.f * fmaxf(s2.y - s1.y * s1.y * invG, 0.f),
                                          sm2.z + 32.f * fmaxf(s2.z - s1.z * s1.z * invG, 0.f), sm2.w + 32.f * fmaxf(s2.w - s1.w * s1.w * invG, 0.f));
            chan_combine4(am, a2, an, gm, g2, 32 * G);
        }
        for (int r = tail_beg + ty; r < b; r += R) chan_combine4(am, a2, an, xp[(size_t)r * Cq], zero, 1);
    }
    s_m[threadIdx.x] = am; s_v[threadIdx.x] = a2; s_n[threadIdx.x] = an;
    __syncthreads();
    if (ty == 0 && quad < Cq) {
        float4 m = make_float4(0.f, 0.f, 0.f, 0.f), v2 = m;
        int n = 0;
        for (int k = 0; k < R; ++k) chan_combine4(m, v2, n, s_m[k * W + tx], s_v[k * W + tx], s_n[k * W + tx]);
        const float inv = 1.0f / (float)max(b - a, 1);
        float* st = stats + (size_t)(seg * nt + t) * 2 * C;
        reinterpret_cast<float4*>(st)[quad] = m;
        reinterpret_cast<float4*>(st + C)[quad] = make_float4(rsqrtf(v2.x * inv + eps), rsqrtf(v2.y * inv + eps),
                                                              rsqrtf(v2.z * inv + eps), rsqrtf(v2.w * inv + eps));
    }
}

// ---- training path: backward of y = LeakyReLU_slope(InstanceNorm(x)) (blocks.py:459-468 + :496-510) -----------------------
// With xhat = (x - mean) * rstd (recovered from y: LeakyReLU is invertible) and dz = dy * act'(xhat):
//   dx = rstd * (dz - mean_rows(dz) - xhat * mean_rows(dz * xhat)).
// Pass 1: per (row chunk, column) partial sums of dz and dz * xhat, fixed order; pass 2: every thread combines the
// partials of its column in chunk order and applies the formula. Deterministic, two launches.
constexpr int NORM_BWD_CHUNKS = 64;

__global__ void __launch_bounds__(256)
norm_bwd_partial_kernel(const float* __restrict__ y, const float* __restrict__ dy, int N, int C, float slope, int rows_per_chunk,
                        float* __restrict__ partial) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const int r0 = blockIdx.y * rows_per_chunk, r1 = min(N, r0 + rows_per_chunk);
    const float inv_slope = 1.0f / slope;
    float s1 = 0.f, s2 = 0.f;
    for (int r = r0; r < r1; ++r) {
        const float yv = y[(size_t)r * C + c], g = dy[(size_t)r * C + c];
        const float xh = yv > 0.f ? yv : yv * inv_slope;
        const float dz = yv > 0.f ? g : g * slope;
        s1 += dz; s2 = fmaf(dz, xh, s2);
    }
    partial[((size_t)blockIdx.y * 2 + 0) * C + c] = s1;
    partial[((size_t)blockIdx.y * 2 + 1) * C + c] = s2;
}

__global__ void __launch_bounds__(256)
norm_bwd_apply_kernel(const float* __restrict__ y, const float* __restrict__ dy, const float* __restrict__ rstd, int N, int C,
                      float slope, int chunks, int rows_per_block, const float* __restrict__ partial, float* __restrict__ dx) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < chunks; ++k) { s1 += partial[((size_t)k * 2 + 0) * C + c]; s2 += partial[((size_t)k * 2 + 1) * C + c]; }
    const float invn = 1.0f / (float)N, m1 = s1 * invn, m2 = s2 * invn, rs = rstd[c], inv_slope = 1.0f / slope;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(N, r0 + rows_per_block);
    for (int r = r0; r < r1; ++r) {
        const float yv = y[(size_t)r * C + c], g = dy[(size_t)r * C + c];
        const float xh = yv > 0.f ? yv : yv * inv_slope;
        const float dz = yv > 0.f ? g : g * slope;
        dx[(size_t)r * C + c] = rs * (dz - m1 - xh * m2);
    }
}

// seg_off[s] = first row of segment s = sum of the lengths of the clouds before cloud s*cps (S+1 entries)
__global__ void seg_offsets_kernel(const int* __restrict__ lens, int B, int cps, int S, int* __restrict__ seg_off) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int acc = 0;
    for (int b = 0; b < B; ++b) {
        if (b % cps == 0) seg_off[b / cps] = acc;
        acc += lens[b];
    }
    seg_off[S] = acc;
}

constexpr int NORM_SEG_MAX_CH = 64;
int g_dbg_skip_d2h = 0;
int g_host_zero_copy = 0;   // aprb_set_option("host_zero_copy"): async host output stored by the last kernel itself (measured slower: 1991 vs 2393 clouds/s)
int g_act_f16 = 1;      // aprb_set_option("act_f16"): aprb_kfe_forward stores normalised activations in fp16 (needs kpconv_f16)
int g_kpconv_f16 = 1;   // aprb_set_option("kpconv_f16"): aprb_kfe_forward runs KPConv with fp16 operands where a block provides them
int g_fuse_stats = 1;   // aprb_set_option("fuse_stats"): aprb_kfe_forward hands GEMM-epilogue group statistics to the norms
int g_gemm_apply = 1;   // aprb_set_option("gemm_apply"): unary2 / shortcut Linears are recomputed with the normalisation in the epilogue

}  // namespace aprb

extern "C" size_t aprb_instnorm_backward_ws_bytes(int C) {
    return C < 0 ? 0 : align256((size_t)NORM_BWD_CHUNKS * 2 * (size_t)C * sizeof(float)) + 256;
}

extern "C" int aprb_instnorm_lrelu_backward(const float* d_y, const float* d_dy, const float* d_rstd, int N, int C, float slope,
                                            float* d_dx, void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(N >= 0 && C >= 1 && slope > 0.f, "need N >= 0, C >= 1 and a positive slope (1 = no activation)");
    if (N == 0) return APRB_OK;
    APRB_REQUIRE(d_y && d_dy && d_rstd && d_dx && d_ws, "null pointer");
    if (ws_bytes < aprb_instnorm_backward_ws_bytes(C)) { set_error("aprb_instnorm_lrelu_backward: workspace too small"); return APRB_ERR_WORKSPACE; }
    const int chunks = min(NORM_BWD_CHUNKS, max(1, cdiv(N, 64)));
    const int rpc = cdiv(N, chunks);
    float* partial = (float*)d_ws;
    APRB_TIMED("norm_bwd_partial_kernel", st, 1, (norm_bwd_partial_kernel<<<dim3(cdiv(C, 256), chunks), 256, 0, st>>>(d_y, d_dy, N, C, slope, rpc, partial)));
    const int ablocks = max(1, min(cdiv(N, 32), 4 * sm_count() / max(1, cdiv(C, 256))));
    const int rpb = cdiv(N, ablocks);
    APRB_TIMED("norm_bwd_apply_kernel", st, 1, (norm_bwd_apply_kernel<<<dim3(cdiv(C, 256), cdiv(N, rpb)), 256, 0, st>>>(d_y, d_dy, d_rstd, N, C, slope, cdiv(N, rpc), rpb, partial, d_dx)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

extern "C" int aprb_segment_offsets(const int32_t* d_lens, int B, int clouds_per_segment, int32_t* d_seg_off, void* stream) {
    APRB_REQUIRE(d_lens && d_seg_off && B >= 1 && clouds_per_segment >= 1, "bad argument");
    const int S = cdiv(B, clouds_per_segment);
    APRB_TIMED("seg_offsets_kernel", (cudaStream_t)stream, 1, (seg_offsets_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_lens, B, clouds_per_segment, S, d_seg_off)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

extern "C" size_t aprb_instnorm_seg_ws_bytes(int N, int C, int S) {
    (void)N;
    if (C < 1 || S < 1) return 0;
    return 2 * align256((size_t)2 * S * NORM_SEG_MAX_CH * C * sizeof(float)) + align256((size_t)4 * S * C * sizeof(float)) +
           align256((size_t)2 * S * (C / 4 + 1) * sizeof(int)) + 256;
}

// Statistics only: d_stats[seg][t][mean | rstd][C] (t = 0 for d_x, 1 for d_x2 when given) from the producers' group
// partials; the tensors themselves are read only at the ragged rows of each segment (aprb_linear_f16_stats_ragged wrote
// exactly those). Consumed by aprb_linear_f16_norm_apply.
extern "C" int aprb_instnorm_seg_stats(const float* d_x, const float* d_x2, int N, int C, const int32_t* d_seg_off, int S,
                                       float eps, const float* d_gstat_x, const float* d_gstat_x2, float* d_stats,
                                       void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(N >= 0 && C >= 4 && C % 4 == 0 && S >= 1, "bad shape (C % 4 == 0)");
    APRB_REQUIRE(S == 1 || d_seg_off, "segment offsets required when S > 1");
    APRB_REQUIRE((d_x2 == nullptr) == (d_gstat_x2 == nullptr), "second tensor and its group statistics go together");
    if (N == 0) return APRB_OK;
    APRB_REQUIRE(d_x && d_gstat_x && d_stats, "null pointer");
    const int nt = d_x2 ? 2 : 1, Cq = C / 4;
    int Wg = 1;
    while (Wg < Cq && Wg < 8) Wg <<= 1;
    APRB_TIMED("norm_seg_groups_kernel", st, 1, (norm_seg_groups_kernel<<<dim3(cdiv(Cq, Wg), S * nt), 256, 0, st>>>(
        d_x, d_x2, d_gstat_x, d_gstat_x2, d_seg_off, N, C, Wg, nt, 0, nt, eps, d_stats)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

extern "C" int aprb_instnorm_lrelu_seg(const float* d_x, int N, int C, const int32_t* d_seg_off, int S, float eps,
                                       float slope, const float* d_residual, int norm_residual, int round_tf32, float* d_y,
                                       void* d_ws, size_t ws_bytes, void* stream) {
    return aprb_instnorm_lrelu_seg_pre(d_x, N, C, d_seg_off, S, eps, slope, d_residual, norm_residual, round_tf32, d_y,
                                       nullptr, nullptr, d_ws, ws_bytes, stream);
}

extern "C" int aprb_instnorm_lrelu_seg_pre(const float* d_x, int N, int C, const int32_t* d_seg_off, int S, float eps,
                                           float slope, const float* d_residual, int norm_residual, int round_tf32,
                                           float* d_y, const float* d_gstat_x, const float* d_gstat_res, void* d_ws,
                                           size_t ws_bytes, void* stream) {
    return aprb_instnorm_lrelu_seg_f16(d_x, N, C, d_seg_off, S, eps, slope, d_residual, 0, norm_residual, round_tf32, d_y, 0,
                                       d_gstat_x, d_gstat_res, d_ws, ws_bytes, stream);
}

extern "C" int aprb_instnorm_lrelu_seg_f16(const float* d_x, int N, int C, const int32_t* d_seg_off, int S, float eps,
                                           float slope, const void* d_residual_v, int residual_is_f16, int norm_residual,
                                           int round_tf32, void* d_y_v, int out_is_f16, const float* d_gstat_x,
                                           const float* d_gstat_res, void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const float* d_residual = (const float*)d_residual_v;
    float* d_y = (float*)d_y_v;
    APRB_REQUIRE(!(residual_is_f16 && norm_residual), "a residual that is standardised itself must be fp32 (a GEMM output)");
    APRB_REQUIRE(!(out_is_f16 && (const void*)d_x == d_y_v), "fp16 output cannot alias the fp32 input");
    APRB_REQUIRE(N >= 0 && C >= 1 && S >= 1, "bad shape");
    APRB_REQUIRE(S == 1 || d_seg_off, "segment offsets required when S > 1");
    if (N == 0) return APRB_OK;
    APRB_REQUIRE(d_x && d_y && d_ws, "null pointer");
    const bool aligned = (((uintptr_t)d_x | (uintptr_t)d_y | (uintptr_t)(d_residual ? d_residual : d_x) |
                           (uintptr_t)(d_gstat_x ? d_gstat_x : d_x) | (uintptr_t)(d_gstat_res ? d_gstat_res : d_x)) & 15) == 0;
    if (C % 4 != 0 || !aligned) { set_error("aprb_instnorm_lrelu_seg: needs C %% 4 == 0 and 16-byte aligned tensors"); return APRB_ERR_UNSUPPORTED; }
    if (ws_bytes < aprb_instnorm_seg_ws_bytes(N, C, S)) { set_error("aprb_instnorm_lrelu_seg: workspace too small"); return APRB_ERR_WORKSPACE; }
    const int nt = (d_residual && norm_residual) ? 2 : 1;
    Carver c(d_ws, ws_bytes);
    float* pmean = c.take<float>((size_t)2 * S * NORM_SEG_MAX_CH * C);
    float* pm2 = c.take<float>((size_t)2 * S * NORM_SEG_MAX_CH * C);
    float* stats = c.take<float>((size_t)4 * S * C);
    int* tickets = c.take<int>((size_t)2 * S * (C / 4 + 1));
    const int Cq = C / 4;
    int W = 1;
    while (W < Cq && W < 256) W <<= 1;
    const int gx = cdiv(Cq, W), sms = sm_count();
    const int rows_seg = max(1, N / S);
    // ---- statistics: from the producer's group partials where given, else one pass over the tensor ----
    const bool pre[2] = {d_gstat_x != nullptr, nt == 2 && d_gstat_res != nullptr};
    for (int t0 = 0; t0 < nt;) {                                      // maximal runs of tensors with the same source
        int t1 = t0 + 1;
        while (t1 < nt && pre[t1] == pre[t0]) ++t1;
        const int tc = t1 - t0;
        if (pre[t0]) {
            int Wg = 1;
            while (Wg < Cq && Wg < 8) Wg <<= 1;                       // narrow column tiles: 256 / Wg lanes over the groups
            APRB_TIMED("norm_seg_groups_kernel", st, 1, (norm_seg_groups_kernel<<<dim3(cdiv(Cq, Wg), S * tc), 256, 0, st>>>(
                d_x, d_residual, d_gstat_x, d_gstat_res, d_seg_off, N, C, Wg, nt, t0, tc, eps, stats)));
        } else {
            // ~4 waves of blocks over all segments, chunks of >= 32 rows
            int ch = min(min(max(1, 4 * sms / (gx * S * tc)), NORM_SEG_MAX_CH), max(1, rows_seg / 32));
            APRB_CUDA_OK(cudaMemsetAsync(tickets, 0, (size_t)S * tc * gx * sizeof(int), st));
            APRB_TIMED("norm_seg_partial_kernel", st, 1, (norm_seg_partial_kernel<<<dim3(gx, ch, S * tc), 256, 0, st>>>(
                d_x, d_residual, d_seg_off, N, C, W, ch, nt, t0, tc, eps, pmean, pm2, tickets, stats)));
        }
        t0 = t1;
    }
    // apply pass: ~6 waves of blocks, >= 16 rows per block
    int rb = min(max(1, 6 * sms / (gx * S)), max(1, rows_seg / 16));
    APRB_TIMED("norm_seg_apply_kernel", st, 1, (norm_seg_apply_kernel<<<dim3(gx, rb, S), 256, 0, st>>>(
        d_x, d_seg_off, N, C, W, nt, stats, d_residual, slope, round_tf32, d_y, residual_is_f16, out_is_f16)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}
