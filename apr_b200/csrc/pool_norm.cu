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
// Pass 1: per (row-chunk, column) Welford partials (count, mean, M2); pass 2: Chan-combine the partials in fixed
// order (deterministic, no float atomics) -> mean, rstd; pass 3: apply.
constexpr int NORM_ROWS_PER_CHUNK = 256;

__global__ void norm_partial_kernel(const float* __restrict__ x, int N, int C, float* __restrict__ pmean,
                                    float* __restrict__ pm2) {
    // grid: (ceil(C/32), chunks); block: (32, 8). Each thread walks rows r = ty, ty+8, ... of its chunk for column c.
    __shared__ float s_mean[8][33], s_m2[8][33];
    __shared__ int s_cnt[8][33];
    int c = blockIdx.x * 32 + threadIdx.x;
    int r0 = blockIdx.y * NORM_ROWS_PER_CHUNK, r1 = min(r0 + NORM_ROWS_PER_CHUNK, N);
    float mean = 0.f, m2 = 0.f;
    int cnt = 0;
    if (c < C) {
        for (int r = r0 + threadIdx.y; r < r1; r += 8) {
            float v = x[(size_t)r * C + c];
            ++cnt;
            float d = v - mean;
            mean += d / (float)cnt;
            m2 += d * (v - mean);
        }
    }
    s_mean[threadIdx.y][threadIdx.x] = mean; s_m2[threadIdx.y][threadIdx.x] = m2; s_cnt[threadIdx.y][threadIdx.x] = cnt;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float am = 0.f, a2 = 0.f;
        int an = 0;
        for (int k = 0; k < 8; ++k) {
            int bn = s_cnt[k][threadIdx.x];
            if (bn == 0) continue;
            float bm = s_mean[k][threadIdx.x], b2 = s_m2[k][threadIdx.x];
            int n = an + bn;
            float d = bm - am;
            am += d * ((float)bn / (float)n);
            a2 += b2 + d * d * ((float)an * (float)bn / (float)n);
            an = n;
        }
        pmean[(size_t)blockIdx.y * C + c] = am;
        pm2[(size_t)blockIdx.y * C + c] = a2;
    }
}

__global__ void norm_finalize_kernel(const float* __restrict__ pmean, const float* __restrict__ pm2, int N, int C,
                                     int chunks, float eps, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float am = 0.f, a2 = 0.f;
    int an = 0;
    for (int k = 0; k < chunks; ++k) {
        int bn = min(NORM_ROWS_PER_CHUNK, N - k * NORM_ROWS_PER_CHUNK);
        float bm = pmean[(size_t)k * C + c], b2 = pm2[(size_t)k * C + c];
        int n = an + bn;
        float d = bm - am;
        am += d * ((float)bn / (float)n);
        a2 += b2 + d * d * ((float)an * (float)bn / (float)n);
        an = n;
    }
    mean_out[c] = am;
    rstd_out[c] = rsqrtf(a2 / (float)N + eps);
}

__global__ void norm_apply_kernel(const float* __restrict__ x, long long total, int C, const float* __restrict__ mean,
                                  const float* __restrict__ rstd, const float* __restrict__ res,
                                  const float* __restrict__ rmean, const float* __restrict__ rrstd, float slope,
                                  float* __restrict__ y) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c = (int)(i % C);
    float v = (x[i] - mean[c]) * rstd[c];
    if (res) {
        float r = res[i];
        if (rmean) r = (r - rmean[c]) * rrstd[c];
        v += r;
    }
    y[i] = v >= 0.f ? v : v * slope;
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

extern "C" size_t aprb_instnorm_ws_bytes(int N, int C) {
    if (N < 0 || C < 0) return 0;
    size_t chunks = (size_t)cdiv(N > 0 ? N : 1, NORM_ROWS_PER_CHUNK);
    return 2 * (2 * align256(chunks * C * sizeof(float)) + 2 * align256(C * sizeof(float))) + 256;
}

static int norm_stats(const float* d_x, int N, int C, float eps, float* pmean, float* pm2, float* mean, float* rstd,
                      cudaStream_t st) {
    int chunks = cdiv(N, NORM_ROWS_PER_CHUNK);
    APRB_TIMED("norm_partial_kernel", st, 1, (norm_partial_kernel<<<dim3(cdiv(C, 32), chunks), dim3(32, 8), 0, st>>>(d_x, N, C, pmean, pm2)));
    APRB_TIMED("norm_finalize_kernel", st, 1, (norm_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>(pmean, pm2, N, C, chunks, eps, mean, rstd)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

extern "C" int aprb_instnorm_lrelu(const float* d_x, int N, int C, float eps, float slope, const float* d_residual,
                                   int norm_residual, float* d_y, void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(N >= 0 && C >= 1, "bad shape");
    if (N == 0) return APRB_OK;
    APRB_REQUIRE(d_x && d_y && d_ws, "null pointer");
    if (ws_bytes < aprb_instnorm_ws_bytes(N, C)) { set_error("aprb_instnorm_lrelu: workspace too small"); return APRB_ERR_WORKSPACE; }
    Carver c(d_ws, ws_bytes);
    size_t chunks = (size_t)cdiv(N, NORM_ROWS_PER_CHUNK);
    float* pmean = c.take<float>(chunks * C); float* pm2 = c.take<float>(chunks * C);
    float* mean = c.take<float>(C); float* rstd = c.take<float>(C);
    float* rpmean = c.take<float>(chunks * C); float* rpm2 = c.take<float>(chunks * C);
    float* rmean = c.take<float>(C); float* rrstd = c.take<float>(C);
    int rc = norm_stats(d_x, N, C, eps, pmean, pm2, mean, rstd, st);
    if (rc) return rc;
    bool nr = d_residual && norm_residual;
    if (nr) { rc = norm_stats(d_residual, N, C, eps, rpmean, rpm2, rmean, rrstd, st); if (rc) return rc; }
    long long total = (long long)N * C;
    APRB_TIMED("norm_apply_kernel", st, 1, (norm_apply_kernel<<<cdiv(total, 256), 256, 0, st>>>(d_x, total, C, mean, rstd, d_residual, nr ? rmean : nullptr,
                                                       nr ? rrstd : nullptr, slope, d_y)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}
