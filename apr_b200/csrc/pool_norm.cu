// pool_norm.cu — K4 (strided max-pool / nearest-upsample gathers) and K6 (InstanceNorm [+residual] [+LeakyReLU]).
//
// K4 replaces max_pool / closest_pool (/root/reference/Predator_APR/models/blocks.py:86-102, :71-83): the shadow
// index Ns selects an all-zero feature row, which takes part in the max exactly like the reference's appended row.
// K6 replaces BatchNormBlock.forward (blocks.py:459-468; nn.InstanceNorm1d over all N rows of the stacked pair,
// eps 1e-5, biased variance, no affine, no running stats) fused with the LeakyReLU(0.1) / residual add that follow it
// in UnaryBlock (:496-510), SimpleBlock (:592) and ResnetBottleneckBlock (:669-681).
#include "common.cuh"

namespace aprb {

template <typename IdxT>
__global__ void max_pool_kernel(const float* __restrict__ x, const IdxT* __restrict__ idx, int ld, int Nq, int Ns,
                                int H, int C, const int* __restrict__ d_width, float* __restrict__ out) {
    // one thread per (query, 4-channel group); C % 4 == 0 path uses float4
    const int groups = C >> 2;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)Nq * groups) return;
    int n = (int)(t / groups), gch = (int)(t % groups);
    int Hn = d_width ? min(H, *d_width) : H;
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

template <typename IdxT>
__global__ void max_pool_scalar_kernel(const float* __restrict__ x, const IdxT* __restrict__ idx, int ld, int Nq,
                                       int Ns, int H, int C, const int* __restrict__ d_width, float* __restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)Nq * C) return;
    int n = (int)(t / C), c = (int)(t % C);
    int Hn = d_width ? min(H, *d_width) : H;
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

extern "C" int aprb_max_pool(const float* d_x, const void* d_idx, int idx_is_i64, int ld_idx, int Nq, int Ns, int H,
                             int C, const int32_t* d_width, float* d_out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(Nq >= 0 && Ns >= 0 && H >= 1 && C >= 1 && ld_idx >= H, "bad shape");
    if (Nq == 0) return APRB_OK;
    APRB_REQUIRE(d_idx && d_out && (d_x || Ns == 0), "null pointer");
    const int T = 256;
    if (C % 4 == 0 && ((uintptr_t)d_x % 16 == 0) && ((uintptr_t)d_out % 16 == 0)) {
        long long total = (long long)Nq * (C / 4);
        if (idx_is_i64) APRB_TIMED("max_pool_kernel", st, 1, (max_pool_kernel<long long><<<cdiv(total, T), T, 0, st>>>(d_x, (const long long*)d_idx, ld_idx, Nq, Ns, H, C, d_width, d_out)));
        else APRB_TIMED("max_pool_kernel", st, 1, (max_pool_kernel<int><<<cdiv(total, T), T, 0, st>>>(d_x, (const int*)d_idx, ld_idx, Nq, Ns, H, C, d_width, d_out)));
    } else {
        long long total = (long long)Nq * C;
        if (idx_is_i64) APRB_TIMED("max_pool_scalar_kernel", st, 1, (max_pool_scalar_kernel<long long><<<cdiv(total, T), T, 0, st>>>(d_x, (const long long*)d_idx, ld_idx, Nq, Ns, H, C, d_width, d_out)));
        else APRB_TIMED("max_pool_scalar_kernel", st, 1, (max_pool_scalar_kernel<int><<<cdiv(total, T), T, 0, st>>>(d_x, (const int*)d_idx, ld_idx, Nq, Ns, H, C, d_width, d_out)));
    }
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
    return 4 * align256((size_t)256 * (C > 0 ? C : 1) * sizeof(float)) + 256;
}

extern "C" int aprb_instnorm_lrelu(const float* d_x, int N, int C, float eps, float slope, const float* d_residual,
                                   int norm_residual, int round_tf32, float* d_y, void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(N >= 0 && C >= 1, "bad shape");
    if (N == 0) return APRB_OK;
    APRB_REQUIRE(d_x && d_y && d_ws, "null pointer");
    if (ws_bytes < aprb_instnorm_ws_bytes(N, C)) { set_error("aprb_instnorm_lrelu: workspace too small"); return APRB_ERR_WORKSPACE; }
    Carver c(d_ws, ws_bytes);
    float* pmean = c.take<float>((size_t)256 * C); float* pm2 = c.take<float>((size_t)256 * C);
    float* rpmean = c.take<float>((size_t)256 * C); float* rpm2 = c.take<float>((size_t)256 * C);
    int rpc, chunks;
    norm_plan(N, C, &rpc, &chunks);
    const dim3 blk(32, 8);
    const bool nr = d_residual && norm_residual;
    APRB_TIMED("norm_partial_kernel", st, 1, (norm_partial_kernel<<<dim3(cdiv(C, 32), chunks), blk, 0, st>>>(d_x, N, C, rpc, pmean, pm2)));
    if (nr) APRB_TIMED("norm_partial_kernel", st, 1, (norm_partial_kernel<<<dim3(cdiv(C, 32), chunks), blk, 0, st>>>(d_residual, N, C, rpc, rpmean, rpm2)));
    int rpb = rpc;   // same row tiling for the apply pass
    APRB_TIMED("norm_apply_kernel", st, 1, (norm_apply_kernel<<<dim3(cdiv(C, 32), cdiv(N, rpb)), blk, 0, st>>>(
        d_x, N, C, chunks, rpc, eps, pmean, pm2, d_residual, nr ? rpmean : nullptr, nr ? rpm2 : nullptr, slope, rpb, round_tf32, d_y)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}
