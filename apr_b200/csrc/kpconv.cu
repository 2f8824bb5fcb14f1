// kpconv.cu — K5: KPConv forward (rigid kernel points, linear influence, sum aggregation).
//
// Replaces KPConv.forward (/root/reference/Predator_APR/models/blocks.py:229-374, non-deformable branch):
//   w[n,k,h]   = max(0, 1 - |s[idx[n,h]] - q[n] - kp[k]| / extent)                  (:269-289, :328-329)
//   wf[n,k,:]  = sum_h w[n,k,h] * x[idx[n,h],:]     (shadow idx == Ns -> zero row)   (:348-354)
//   out[n,:]   = (sum_k wf[n,k,:] @ W[k]) / max(1, #{h : sum_c x[idx[n,h],c] > 0})   (:361-372)
// Stage A+B (kp_weighted_kernel): one warp per query; influence weights are produced in shared memory, then the
// neighbour feature rows are streamed once per 128-channel slab and accumulated only for the kernel points whose
// influence is non-zero (a neighbour is inside the extent of ~1.4 of the 15 kernel points on average; the test is
// warp-uniform, so skipping costs no divergence). wf is written as the [Nq, K*Cin] A operand of stage C.
// Stage C: [Nq, K*Cin] x [K*Cin, Cout] contraction with the 1/neighbor_num row scale in the epilogue —
// fp32 CUDA-core tiles here (mode 1); the tcgen05 TF32 path lives in gemm_tcgen05.cu (mode 2).
#include "kpconv_common.cuh"
#include <cuda_fp16.h>

namespace aprb {

int gemm_tf32_rowscale(const float* d_A, const float* d_Bt, int M, int N, int K, const float* d_rowscale, float* d_C,
                       void* d_ws, size_t ws_bytes, cudaStream_t st, float* d_gstat, int* stats_written);  // gemm_tcgen05.cu
size_t gemm_tf32_ws_bytes(int M, int N);
bool gemm_f16_supported(int M, int N, int K);
int gemm_f16_rowscale(const void* d_A, const void* d_Bt, int M, int N, int K, const float* d_rowscale, float* d_C,
                      cudaStream_t st, float* d_gstat, int* stats_written);
bool kpconv_fused_supported(int H, int K, int Cin, int Cout, long long Ns);   // kpconv_fused.cu
int kpconv_fused_run(const float* d_q, const float4* s4, const void* d_idx, int idx_is_i64, int ld, const float* d_x,
                     const float* d_kp, const float* d_wprep, float extent, int Nq, int Ns, int H, int K, int Cin, int Cout,
                     float* d_out, float* d_gstat, cudaStream_t st);
bool gemm_tf32_supported(int M, int N, int K);
void set_gemm_label(const char* label);                                           // gemm_tcgen05.cu: timer label of this thread's next GEMMs
bool kpconv_tc_supported(int H, int K, int Cin, long long Ns);                 // kpconv_tc.cu
int kpconv_tc_run(const float* d_q, const float4* s4, const int* d_idx, int ld, const void* d_x16, const float* d_kp,
                  float extent, int Nq, int Ns, int H, int K, int Cin, void* d_wf16, float* d_inv_nn, cudaStream_t st);

int g_kpw_version = 4;       // aprb_set_option("kpw_version"): 3 = per-kernel-point tables, 4 = CSR lists + lane groups
int g_kpw_fh = 1;            // aprb_set_option("kpw_fh"): fp16 rows through the FHFMA kernel (v5); 0 = v4, 2 = Cin 64 with 2 rows per warp
int g_kpconv_chunk_mb = 0;   // aprb_set_option("kpconv_chunk_mb"): L2-sized row chunks of the tensor path (0 = off)

// flag[s] = 1 iff sum_c x[s,c] > 0 (one warp per support row; fixed reduction order), and the packed support record
// s4[s] = (x, y, z, flag) — with C == 1 (x, y, z, feature) —: the gather kernels fetch a neighbour's position and flag with ONE 128-bit load — a
// divergent warp load costs one L1 wavefront per distinct line whatever its width, and the three coordinate loads plus
// the flag byte were 256 of the ~470 LSU wavefronts per query (ncu, round 1).
__global__ void rowsum_pos_kernel(const float* __restrict__ x, const float* __restrict__ pts, int Ns, int C,
                                  unsigned char* __restrict__ flag, float4* __restrict__ s4, int x16) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= Ns) return;
    float s = 0.f;
    if (x16) {
        const __half* xh = reinterpret_cast<const __half*>(x);
        for (int c = lane; c < C; c += 32) s += __half2float(xh[(size_t)row * C + c]);
    } else {
        for (int c = lane; c < C; c += 32) s += x[(size_t)row * C + c];
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) {
        flag[row] = s > 0.f ? 1 : 0;
        s4[row] = make_float4(pts[3 * (size_t)row], pts[3 * (size_t)row + 1], pts[3 * (size_t)row + 2],
                              C == 1 ? s : (s > 0.f ? 1.f : 0.f));   // Cin == 1: the record carries the feature itself
    }
}

// The same for fp16 feature rows of C = 8 * LPR channels (LPR lanes per row, 16-byte loads, 32 / LPR rows per warp): the
// warp-per-row form above moves 128 bytes per warp and load at C = 64 and ran at 0.7 TB/s (11 launches, 0.26 ms per call).
template <int LPR>
__global__ void __launch_bounds__(256)
rowsum_pos16_kernel(const __half* __restrict__ x, const float* __restrict__ pts, int Ns, unsigned char* __restrict__ flag,
                    float4* __restrict__ s4) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, li = lane % LPR;
    const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + sub;
    float s = 0.f;
    if (row < Ns) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (size_t)row * (LPR * 8)) + li);
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); s += f.x + f.y; }
    }
#pragma unroll
    for (int d = LPR / 2; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (row < Ns && li == 0) {
        flag[row] = s > 0.f ? 1 : 0;
        s4[row] = make_float4(pts[3 * (size_t)row], pts[3 * (size_t)row + 1], pts[3 * (size_t)row + 2], s > 0.f ? 1.f : 0.f);
    }
}

// flag / packed support records of a feature table, by the kernel that fits it
static void launch_rowsum_pos(const float* d_x, const float* d_s, int Ns, int C, unsigned char* flag, float4* s4, int x16, cudaStream_t st) {
    if (x16 && ((uintptr_t)d_x & 15) == 0 && (C == 64 || C == 128 || C == 256)) {
        const __half* xh = reinterpret_cast<const __half*>(d_x);
        if (C == 64) APRB_TIMED("rowsum_pos_kernel", st, 1, (rowsum_pos16_kernel<8><<<cdiv(Ns, 32), 256, 0, st>>>(xh, d_s, Ns, flag, s4)));
        else if (C == 128) APRB_TIMED("rowsum_pos_kernel", st, 1, (rowsum_pos16_kernel<16><<<cdiv(Ns, 16), 256, 0, st>>>(xh, d_s, Ns, flag, s4)));
        else APRB_TIMED("rowsum_pos_kernel", st, 1, (rowsum_pos16_kernel<32><<<cdiv(Ns, 8), 256, 0, st>>>(xh, d_s, Ns, flag, s4)));
        return;
    }
    APRB_TIMED("rowsum_pos_kernel", st, 1, (rowsum_pos_kernel<<<cdiv(Ns, 8), 256, 0, st>>>(d_x, d_s, Ns, C, flag, s4, x16)));
}

template <int VEC> struct Vec;
template <> struct Vec<1> { float v[1]; };
template <> struct __align__(8) Vec<2> { float v[2]; };
template <> struct __align__(16) Vec<4> { float v[4]; };

// One warp per query (stage A+B of KPConv).
// Phase 1 (lanes = neighbours): relative positions, influence w = max(0, 1 - d/extent) against the K kernel points,
//   neighbor_num, and — per kernel point — a compacted list of (support index, w) of the neighbours it influences
//   (ballot compaction; ~5 of H=57 neighbours per kernel point). Dynamic smem per warp: K_MAX lists x Hp x 8 bytes.
// Phase 2 (lanes = channels): for each kernel point, stream the feature rows of its list (NU rows in flight) and
//   accumulate w * x into ONE VEC*NJ-wide register tile, then store that kernel point's slice of wf. The kernel point is
//   a plain loop variable — no per-neighbour dispatch — and a row that influences several kernel points (1.4 on average)
//   is simply re-read (L1/L2 hit).
template <typename IdxT, int VEC, int NJ, bool ROUND_TF32>
__global__ void __launch_bounds__(128)
kp_weighted_kernel(const float* __restrict__ q, const float* __restrict__ s, const IdxT* __restrict__ idx, int ld,
                   const float* __restrict__ x, const float* __restrict__ kp, const unsigned char* __restrict__ posflag,
                   float extent, int Nq, int Ns, int H, int K, int Cin, float* __restrict__ wf,
                   float* __restrict__ inv_nn) {
    extern __shared__ float s_dyn[];
    __shared__ float s_kp[KP_MAX_K * 3];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int Hp = (H + 31) & ~31;
    if (threadIdx.x < K * 3) s_kp[threadIdx.x] = kp[threadIdx.x];
    __syncthreads();
    const int n = blockIdx.x * wpb + wib;
    if (n >= Nq) return;
    int* s_si = reinterpret_cast<int*>(s_dyn) + (size_t)wib * KP_MAX_K * Hp * 2;   // [K_MAX][Hp] support index
    float* s_lw = reinterpret_cast<float*>(s_si + (size_t)KP_MAX_K * Hp);          // [K_MAX][Hp] influence weight
    int* s_cnt = reinterpret_cast<int*>(s_dyn) + (size_t)wpb * KP_MAX_K * Hp * 2 + wib * KP_MAX_K;   // [K_MAX] list lengths
    const float qx = q[3 * (size_t)n], qy = q[3 * (size_t)n + 1], qz = q[3 * (size_t)n + 2];
    const float ext2 = extent * extent, inv_ext = 1.0f / extent;

    int nn = 0;
    int cnt[KP_MAX_K];
#pragma unroll
    for (int k = 0; k < KP_MAX_K; ++k) cnt[k] = 0;
    for (int h0 = 0; h0 < Hp; h0 += 32) {
        const int h = h0 + lane;
        int si = Ns;
        if (h < H) {
            const long long v = (long long)idx[(size_t)n * ld + h];
            si = (v >= 0 && v < Ns) ? (int)v : Ns;
        }
        float rx = 0.f, ry = 0.f, rz = 0.f;
        const bool valid = si < Ns;
        if (valid) {
            nn += posflag[si];
            rx = s[3 * (size_t)si] - qx; ry = s[3 * (size_t)si + 1] - qy; rz = s[3 * (size_t)si + 2] - qz;
        }
        if (!__any_sync(0xffffffffu, valid)) continue;
#pragma unroll
        for (int k = 0; k < KP_MAX_K; ++k) {
            if (k < K) {
                const float ddx = rx - s_kp[3 * k], ddy = ry - s_kp[3 * k + 1], ddz = rz - s_kp[3 * k + 2];
                const float d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                float w = 0.f;
                if (valid && d2 < ext2) w = 1.0f - sqrtf(d2) * inv_ext;
                const bool in = w > 0.f;
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (in) {
                    const int pos = cnt[k] + __popc(m & ((1u << lane) - 1));
                    s_si[k * Hp + pos] = si * Cin;              // element offset of the support's feature row
                    s_lw[k * Hp + pos] = w;
                }
                cnt[k] += __popc(m);
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) nn += __shfl_xor_sync(0xffffffffu, nn, d);
    if (lane == 0 && blockIdx.y == 0) inv_nn[n] = 1.0f / (float)max(nn, 1);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < KP_MAX_K; ++k) s_cnt[k] = cnt[k];
    }
    __syncwarp();

    float* wrow = wf + (size_t)n * K * Cin;
    constexpr int SLAB = 32 * VEC * NJ;
    constexpr int NU = (VEC * NJ >= 8) ? 2 : 4;                   // rows in flight
    for (int c0 = blockIdx.y * SLAB; c0 < Cin; c0 += gridDim.y * SLAB) {
        const float* xs = x + c0 + lane * VEC;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {                                // runtime loop: small code, stays in the I-cache
            {
                float acc[NJ][VEC];
#pragma unroll
                for (int j = 0; j < NJ; ++j)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[j][v] = 0.f;
                const int ck = s_cnt[k];
                const int* lsi = s_si + k * Hp;
                const float* lw = s_lw + k * Hp;
#pragma unroll 1
                for (int e0 = 0; e0 < ck; e0 += NU) {
                    Vec<VEC> xr[NU][NJ];
                    float w[NU];
#pragma unroll
                    for (int u = 0; u < NU; ++u) {
                        const bool on = e0 + u < ck;
                        w[u] = on ? lw[e0 + u] : 0.f;
                        const float* row = xs + (on ? lsi[e0 + u] : lsi[e0]);
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            if (c0 + lane * VEC + j * 32 * VEC < Cin) xr[u][j] = *reinterpret_cast<const Vec<VEC>*>(row + j * 32 * VEC);
                            else {
#pragma unroll
                                for (int v = 0; v < VEC; ++v) xr[u][j].v[v] = 0.f;
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < NU; ++u)
#pragma unroll
                        for (int j = 0; j < NJ; ++j)
#pragma unroll
                            for (int v = 0; v < VEC; ++v) acc[j][v] = fmaf(w[u], xr[u][j].v[v], acc[j][v]);
                }
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int c = c0 + lane * VEC + j * 32 * VEC;
                    if (c < Cin) {
                        Vec<VEC> o;
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            float t = acc[j][v];
                            if (ROUND_TF32) {
                                unsigned u;
                                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(t));
                                t = __uint_as_float(u);
                            }
                            o.v[v] = t;
                        }
                        *reinterpret_cast<Vec<VEC>*>(wrow + (size_t)k * Cin + c) = o;
                    }
                }
            }
        }
    }
}

// ---- v4: compact per-row CSR influence lists + lane-group streaming ---------------------------------------------------
// Same contract as kp_weighted_kernel (Cin % 4 == 0, 16-byte aligned x / wf, H <= 32*NH). What changes is where the
// instructions go (the v3 kernel was instruction-issue-bound, ~2700 warp instructions per query at Cin = 64):
//  * phase 1 runs kernel-point-major, so the (element offset, weight) pairs of kernel point k land contiguously after
//    those of k-1: one CSR row of KPW_ECAP entries + K+1 offsets (1.3 KB) replaces the K x Hp table (8 KB) — 6x the
//    resident warps; the kernel point is one LDS.128, the square root is sqrt.approx (1 ulp; the weight is linear in it);
//  * phase 2 splits the warp into 32/LG lane groups that stream DIFFERENT rows: a group of LG lanes covers LG*4*NV
//    channels with 128-bit loads, so one LDS.64 + one address + NV LDG.128 + 4*NV FFMA serve 32/LG list entries;
//    lists shorter than the longest of the warp's rows are padded with weight-0 reads of row 0.
// A row whose list would exceed KPW_ECAP entries (kernel points much closer together than the extent) is evaluated
// directly by the whole warp (kp_direct_row): ballot over the neighbours, shuffle-broadcast of (offset, weight).
constexpr int KPW_ECAP = 160;
constexpr int KPW_SLOT_BYTES = KPW_ECAP * 8 + 80;   // int2 ent[ECAP]; int off[K_MAX+1] (+pad)

// fp16 form of the weighted tile (tensor path with fp16 operands: same 10-bit mantissa as TF32, half the bytes)
__device__ __forceinline__ void store_half4(__half* dst, const float4 v) {
    uint2 u;
    u.x = pack_half2_sat(v.x, v.y); u.y = pack_half2_sat(v.z, v.w);
    *reinterpret_cast<uint2*>(dst) = u;
}

// 4 consecutive channels of a feature row: fp32 (16 bytes) or fp16 (8 bytes, widened; activations stored in fp16 are
// TF32-rounded values, so the widening is exact)
template <bool X16>
__device__ __forceinline__ float4 ld_feat4(const char* p) {
    if (X16) {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
        const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    return __ldg(reinterpret_cast<const float4*>(p));
}

// Overflow path: one row, whole warp, lanes = 4 channels of a 128-channel slab; no lists.
template <typename IdxT, int NH, bool ROUND_TF32, bool X16>
__device__ __noinline__ void kp_direct_row(const float* __restrict__ q, const float4* __restrict__ s4, const IdxT* __restrict__ idx,
                                           int ld, const float* __restrict__ x, const float4* s_kp,
                                           float ext2, float inv_ext, int n, int Ns,
                                           int H, int K, int Cin, float* __restrict__ wrow, int out16, int lane) {
    RowGeom<NH> g;
    load_row_geom<IdxT, NH>(q, s4, idx, ld, n, Ns, H, Cin * (X16 ? 2 : 4), lane, g);
    for (int c0 = 0; c0 < Cin; c0 += 128) {
        const int c = c0 + lane * 4;
        const bool cok = c < Cin;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const float4 kpk = s_kp[k];
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < NH; ++j) {
                if (!g.any[j]) continue;
                bool in;
                const float w = influence<NH>(g, j, kpk, ext2, inv_ext, in);
                unsigned m = __ballot_sync(0xffffffffu, in);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float wb = __shfl_sync(0xffffffffu, w, src);
                    const int sib = __shfl_sync(0xffffffffu, g.sio[j], src);
                    if (cok) {
                        const float4 xr = ld_feat4<X16>(reinterpret_cast<const char*>(x) + (unsigned)sib + (unsigned)c * (X16 ? 2u : 4u));
                        acc.x = fmaf(wb, xr.x, acc.x); acc.y = fmaf(wb, xr.y, acc.y);
                        acc.z = fmaf(wb, xr.z, acc.z); acc.w = fmaf(wb, xr.w, acc.w);
                    }
                }
            }
            if (cok) {
                if (out16) {
                    store_half4(reinterpret_cast<__half*>(wrow) + (size_t)k * Cin + c, acc);
                } else {
                    if (ROUND_TF32) { acc.x = round_tf32(acc.x); acc.y = round_tf32(acc.y); acc.z = round_tf32(acc.z); acc.w = round_tf32(acc.w); }
                    *reinterpret_cast<float4*>(wrow + (size_t)k * Cin + c) = acc;
                }
            }
        }
    }
}

template <typename IdxT, int LG, int NV, int NH, bool FULL, bool ROUND_TF32, bool X16>
__global__ void __launch_bounds__(128)
kp_weighted4_kernel(const float* __restrict__ q, const float4* __restrict__ s4, const IdxT* __restrict__ idx, int ld,
                    const float* __restrict__ x, const float* __restrict__ kp,
                    float extent, int Nq, int Ns, int H, int K, int Cin, float* __restrict__ wf,
                    float* __restrict__ inv_nn, int out16) {
    constexpr int RP = 32 / LG;                  // rows streamed in parallel by one warp
    constexpr int CH = LG * 4 * NV;              // channels per pass
    constexpr int NU = 2;                        // list entries in flight per lane group (lists average ~3 entries)
    extern __shared__ __align__(16) unsigned char s_rows[];
    __shared__ float4 s_kp[KP_MAX_K];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    if (threadIdx.x < K) s_kp[threadIdx.x] = make_float4(kp[3 * threadIdx.x], kp[3 * threadIdx.x + 1], kp[3 * threadIdx.x + 2], 0.f);
    __syncthreads();
    const int row0 = (blockIdx.x * wpb + wib) * RP;
    if (row0 >= Nq) return;
    const float ext2 = extent * extent, inv_ext = 1.0f / extent;
    unsigned char* wslots = s_rows + (size_t)wib * RP * KPW_SLOT_BYTES;
    const unsigned ltmask = (1u << lane) - 1u;

    // ---- phase 1: CSR influence lists of the warp's RP rows (whole warp per row, lanes = neighbours) ----
#pragma unroll 1
    for (int r = 0; r < RP; ++r) {
        int2* ent = reinterpret_cast<int2*>(wslots + (size_t)r * KPW_SLOT_BYTES);
        int* off = reinterpret_cast<int*>(ent + KPW_ECAP);
        const int n = row0 + r;
        if (n >= Nq) {
            if (lane <= KP_MAX_K) off[lane] = 0;
            continue;
        }
        RowGeom<NH> g;
        const int nn = load_row_geom<IdxT, NH>(q, s4, idx, ld, n, Ns, H, Cin * (X16 ? 2 : 4), lane, g);
        if (lane == 0) inv_nn[n] = 1.0f / (float)max(nn, 1);
        build_row_list<NH, KPW_ECAP>(g, s_kp, K, ext2, inv_ext, ent, off, lane);
    }
    __syncwarp();

    // ---- phase 2: lane group grp streams row row0 + grp ----
    // Each lane group runs its OWN trip count (no warp-level primitive inside the loops): shorter lists simply leave
    // their lanes masked off while the longest one finishes.
    const int grp = lane / LG, lg = lane % LG;
    const int n = row0 + grp;
    const int2* ent = reinterpret_cast<const int2*>(wslots + (size_t)grp * KPW_SLOT_BYTES);
    const int* off = reinterpret_cast<const int*>(ent + KPW_ECAP);
    const bool ovf = off[KP_MAX_K] > KPW_ECAP;
    const bool active = n < Nq && !ovf;
    const char* xb = reinterpret_cast<const char*>(x);
    for (int c0 = 0; c0 < Cin; c0 += CH) {
        const int c = c0 + lg * 4;
        bool cok[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) cok[j] = FULL || (c + j * LG * 4 < Cin);
        constexpr int ES = X16 ? 2 : 4;                              // bytes per stored feature
        const char* xbc = xb + (size_t)c * ES;
        size_t wo = (size_t)n * K * Cin + c;                       // element offset of (row n, kernel point k, channel c)
        int end = active ? off[0] : 0;
#pragma unroll 1
        for (int k = 0; k < K; ++k, wo += Cin) {
            const int beg = end;
            end = active ? off[k + 1] : 0;
            float4 acc[NV];
#pragma unroll
            for (int j = 0; j < NV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            int e = beg;
#pragma unroll 1
            for (; e + NU <= end; e += NU) {
                int2 en[NU];
                float4 xr[NU][NV];
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    en[u] = ent[e + u];
                    const char* xe = xbc + (unsigned)en[u].x;
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        xr[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (cok[j]) xr[u][j] = ld_feat4<X16>(xe + j * LG * 4 * ES);
                    }
                }
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const float w = __int_as_float(en[u].y);
#pragma unroll
                    for (int j = 0; j < NV; ++j) {
                        acc[j].x = fmaf(w, xr[u][j].x, acc[j].x); acc[j].y = fmaf(w, xr[u][j].y, acc[j].y);
                        acc[j].z = fmaf(w, xr[u][j].z, acc[j].z); acc[j].w = fmaf(w, xr[u][j].w, acc[j].w);
                    }
                }
            }
#pragma unroll 1
            for (; e < end; ++e) {
                const int2 en = ent[e];
                const char* xe = xbc + (unsigned)en.x;
                const float w = __int_as_float(en.y);
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    if (cok[j]) {
                        const float4 xr = ld_feat4<X16>(xe + j * LG * 4 * ES);
                        acc[j].x = fmaf(w, xr.x, acc[j].x); acc[j].y = fmaf(w, xr.y, acc[j].y);
                        acc[j].z = fmaf(w, xr.z, acc[j].z); acc[j].w = fmaf(w, xr.w, acc[j].w);
                    }
                }
            }
            if (active) {
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    if (cok[j]) {
                        float4 o4 = acc[j];
                        if (out16) {
                            store_half4(reinterpret_cast<__half*>(wf) + wo + j * LG * 4, o4);
                        } else {
                            if (ROUND_TF32) { o4.x = pre_round_tf32(o4.x); o4.y = pre_round_tf32(o4.y); o4.z = pre_round_tf32(o4.z); o4.w = pre_round_tf32(o4.w); }
                            *reinterpret_cast<float4*>(wf + wo + j * LG * 4) = o4;
                        }
                    }
                }
            }
        }
    }
    __syncwarp();
    // ---- rows whose lists overflowed: direct evaluation by the whole warp ----
    if (__any_sync(0xffffffffu, ovf && n < Nq)) {
#pragma unroll 1
        for (int r = 0; r < RP; ++r) {
            const int ovr = __shfl_sync(0xffffffffu, (int)ovf, r * LG);
            if (ovr && row0 + r < Nq)
                kp_direct_row<IdxT, NH, ROUND_TF32, X16>(q, s4, idx, ld, x, s_kp, ext2, inv_ext, row0 + r, Ns, H, K, Cin,
                                                    out16 ? reinterpret_cast<float*>(reinterpret_cast<__half*>(wf) + (size_t)(row0 + r) * K * Cin)
                                                          : wf + (size_t)(row0 + r) * K * Cin, out16, lane);
        }
    }
}

// ---- v5 (fp16 features in, fp16 weighted tile out): the v4 structure with FHFMA ------------------------------------------
// sm_100a has a mixed-precision FMA, fma.rn.f32.f16 (SASS FHFMA: fp16 x fp16 product, exact, added to an fp32 accumulator
// with one rounding; it reads either half of a packed register). With the influence weight kept as fp16 bits in the list
// entry the inner loop needs no widening of the feature row at all: per list entry LDS.64 + address + ONE 128-bit load
// + 8 FHFMA for 8 channels, against LDS.64 + address + 2 LDG.64 + 8 conversions + 8 FFMA in v4 (phase 2 was 47 % of the
// kernel's issued instructions). What it costs: the weight is rounded to an 11-bit significand (as the tcgen05 weighting
// kernel does, kpconv_tc.cu); the accumulation stays fp32.
__device__ __forceinline__ float fhfma(unsigned short a, unsigned short b, float c) {
    asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(c) : "h"(a), "h"(b));
    return c;
}
template <int NW> struct HWords { uint32_t r[NW]; };
template <int NW> __device__ __forceinline__ HWords<NW> ld_hwords(const char* p);
template <> __device__ __forceinline__ HWords<2> ld_hwords<2>(const char* p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p)); HWords<2> h; h.r[0] = u.x; h.r[1] = u.y; return h;
}
template <> __device__ __forceinline__ HWords<4> ld_hwords<4>(const char* p) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p)); HWords<4> h; h.r[0] = u.x; h.r[1] = u.y; h.r[2] = u.z; h.r[3] = u.w; return h;
}

// Overflow path of v5: as kp_direct_row, fp16 rows, weight rounded to fp16 like the list entries.
template <int NH>
__device__ __noinline__ void kp_direct_row_h(const float* __restrict__ q, const float4* __restrict__ s4, const int* __restrict__ idx,
                                             int ld, const __half* __restrict__ x, const float4* s_kp, float ext2, float inv_ext,
                                             int n, int Ns, int H, int K, int Cin, __half* __restrict__ wrow, int lane) {
    RowGeom<NH> g;
    load_row_geom<int, NH>(q, s4, idx, ld, n, Ns, H, Cin * 2, lane, g);
    for (int c0 = 0; c0 < Cin; c0 += 128) {
        const int c = c0 + lane * 4;
        const bool cok = c < Cin;
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const float4 kpk = s_kp[k];
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < NH; ++j) {
                if (!g.any[j]) continue;
                bool in;
                const float w = influence<NH>(g, j, kpk, ext2, inv_ext, in);
                unsigned m = __ballot_sync(0xffffffffu, in);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const unsigned short wb = (unsigned short)__shfl_sync(0xffffffffu, (int)__half_as_ushort(__float2half_rn(w)), src);
                    const int sib = __shfl_sync(0xffffffffu, g.sio[j], src);
                    if (cok) {
                        const HWords<2> xr = ld_hwords<2>(reinterpret_cast<const char*>(x) + (unsigned)sib + (unsigned)c * 2u);
                        acc.x = fhfma(wb, (unsigned short)(xr.r[0] & 0xffffu), acc.x); acc.y = fhfma(wb, (unsigned short)(xr.r[0] >> 16), acc.y);
                        acc.z = fhfma(wb, (unsigned short)(xr.r[1] & 0xffffu), acc.z); acc.w = fhfma(wb, (unsigned short)(xr.r[1] >> 16), acc.w);
                    }
                }
            }
            if (cok) store_half4(wrow + (size_t)k * Cin + c, acc);
        }
    }
}

// LG lanes per row, CPL (4 or 8) consecutive channels per lane and load, NV loads per lane: one pass covers LG*CPL*NV channels.
template <int LG, int CPL, int NV, int NH>
__global__ void __launch_bounds__(128)
kp_weighted_h_kernel(const float* __restrict__ q, const float4* __restrict__ s4, const int* __restrict__ idx, int ld,
                     const __half* __restrict__ x, const float* __restrict__ kp, float extent, int Nq, int Ns, int H, int K,
                     int Cin, __half* __restrict__ wf, float* __restrict__ inv_nn) {
    constexpr int RP = 32 / LG;
    constexpr int NW = CPL / 2;
    constexpr int CH = LG * CPL * NV;
    constexpr int NU = 2;
    extern __shared__ __align__(16) unsigned char s_rows[];
    __shared__ float4 s_kp[KP_MAX_K];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    if (threadIdx.x < K) s_kp[threadIdx.x] = make_float4(kp[3 * threadIdx.x], kp[3 * threadIdx.x + 1], kp[3 * threadIdx.x + 2], 0.f);
    __syncthreads();
    const int row0 = (blockIdx.x * wpb + wib) * RP;
    if (row0 >= Nq) return;
    const float ext2 = extent * extent, inv_ext = 1.0f / extent;
    unsigned char* wslots = s_rows + (size_t)wib * RP * KPW_SLOT_BYTES;

    // ---- phase 1: CSR influence lists of the warp's RP rows (whole warp per row, lanes = neighbours) ----
#pragma unroll 1
    for (int r = 0; r < RP; ++r) {
        int2* ent = reinterpret_cast<int2*>(wslots + (size_t)r * KPW_SLOT_BYTES);
        int* off = reinterpret_cast<int*>(ent + KPW_ECAP);
        const int n = row0 + r;
        if (n >= Nq) {
            if (lane <= KP_MAX_K) off[lane] = 0;
            continue;
        }
        RowGeom<NH> g;
        const int nn = load_row_geom<int, NH>(q, s4, idx, ld, n, Ns, H, Cin * 2, lane, g);
        if (lane == 0) inv_nn[n] = 1.0f / (float)max(nn, 1);
        build_row_list<NH, KPW_ECAP, true>(g, s_kp, K, ext2, inv_ext, ent, off, lane);
    }
    __syncwarp();

    // ---- phase 2: lane group grp streams row row0 + grp (own trip counts per group, no warp primitive inside) ----
    const int grp = lane / LG, lg = lane % LG;
    const int n = row0 + grp;
    const int2* ent = reinterpret_cast<const int2*>(wslots + (size_t)grp * KPW_SLOT_BYTES);
    const int* off = reinterpret_cast<const int*>(ent + KPW_ECAP);
    const bool ovf = off[KP_MAX_K] > KPW_ECAP;
    const bool active = n < Nq && !ovf;
    const char* xb = reinterpret_cast<const char*>(x);
    for (int c0 = 0; c0 < Cin; c0 += CH) {
        const int c = c0 + lg * CPL;
        const char* xbc = xb + (size_t)c * 2;
        size_t wo = (size_t)n * K * Cin + c;
        int end = active ? off[0] : 0;
#pragma unroll 1
        for (int k = 0; k < K; ++k, wo += Cin) {
            const int beg = end;
            end = active ? off[k + 1] : 0;
            float acc[NV][CPL];
#pragma unroll
            for (int j = 0; j < NV; ++j)
#pragma unroll
                for (int i = 0; i < CPL; ++i) acc[j][i] = 0.f;
            int e = beg;
#pragma unroll 1
            for (; e + NU <= end; e += NU) {
                int2 en[NU];
                HWords<NW> xr[NU][NV];
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    en[u] = ent[e + u];
                    const char* xe = xbc + (unsigned)en[u].x;
#pragma unroll
                    for (int j = 0; j < NV; ++j) xr[u][j] = ld_hwords<NW>(xe + j * LG * CPL * 2);
                }
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    const unsigned short w = (unsigned short)en[u].y;
#pragma unroll
                    for (int j = 0; j < NV; ++j)
#pragma unroll
                        for (int i = 0; i < NW; ++i) {
                            acc[j][2 * i] = fhfma(w, (unsigned short)(xr[u][j].r[i] & 0xffffu), acc[j][2 * i]);
                            acc[j][2 * i + 1] = fhfma(w, (unsigned short)(xr[u][j].r[i] >> 16), acc[j][2 * i + 1]);
                        }
                }
            }
#pragma unroll 1
            for (; e < end; ++e) {
                const int2 en = ent[e];
                const char* xe = xbc + (unsigned)en.x;
                const unsigned short w = (unsigned short)en.y;
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    const HWords<NW> xr = ld_hwords<NW>(xe + j * LG * CPL * 2);
#pragma unroll
                    for (int i = 0; i < NW; ++i) {
                        acc[j][2 * i] = fhfma(w, (unsigned short)(xr.r[i] & 0xffffu), acc[j][2 * i]);
                        acc[j][2 * i + 1] = fhfma(w, (unsigned short)(xr.r[i] >> 16), acc[j][2 * i + 1]);
                    }
                }
            }
            if (active) {
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    uint32_t o[NW];
#pragma unroll
                    for (int i = 0; i < NW; ++i) o[i] = pack_half2_sat(acc[j][2 * i], acc[j][2 * i + 1]);
                    __half* dst = wf + wo + j * LG * CPL;
                    if (NW == 2) *reinterpret_cast<uint2*>(dst) = make_uint2(o[0], o[1]);
                    else *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[NW - 2], o[NW - 1]);
                }
            }
        }
    }
    __syncwarp();
    if (__any_sync(0xffffffffu, ovf && n < Nq)) {
#pragma unroll 1
        for (int r = 0; r < RP; ++r) {
            const int ovr = __shfl_sync(0xffffffffu, (int)ovf, r * LG);
            if (ovr && row0 + r < Nq)
                kp_direct_row_h<NH>(q, s4, idx, ld, x, s_kp, ext2, inv_ext, row0 + r, Ns, H, K, Cin, wf + (size_t)(row0 + r) * K * Cin, lane);
        }
    }
}

// (Round 2 also measured v5 with the list consumed as ONE stream — entries tagged with "last of kernel point k", 4 entries
// and their rows in flight across kernel-point boundaries: bit-identical output, same time within 3 % on all seven encoder
// shapes (gpurun log r02c), so phase 2 is not latency-bound either; the per-kernel-point form stays.
// And v6, list building with lane = (kernel point, neighbour parity) walking the row's compacted neighbours and appending
// to its own fixed slot — no ballots, no prefix sums, ~500 instead of ~830 instructions per row on paper: 1.4-2x SLOWER
// (583 vs 425 us at L0). Real kernel points with KP_extent 2.0 put up to 27 neighbours in one extent, so slots of 24
// entries per kernel point are needed (1.9 KB per row: occupancy), and the walk is a serial, branchy latency chain where
// v4/v5's 30 evaluations per row are independent and fully unrolled. Removed; numbers in DESIGN.md 4.4d.)

// (Round 1 also measured this stage on mma.sync — dense per-query 16 x H x Cin products, three variants. On B200 every
// legacy HMMA.1688 is charged ~4 LSU data-pipe wavefronts, so that path is LSU-bound at the speed of this kernel or
// worse: profiles/r01_kp_weighted_variants.txt; the code is in the history at commit "KPConv weighting on mma.sync".)

// Cin == 1 (the first encoder block: one scalar feature per point): lanes = neighbours, the 15 per-kernel-point sums
// are reduced across the warp with shuffles; no shared memory, no second pass.
template <typename IdxT>
__global__ void __launch_bounds__(128)
kp_weighted_c1_kernel(const float* __restrict__ q, const float4* __restrict__ s4, const IdxT* __restrict__ idx, int ld,
                      const float* __restrict__ kp, float extent, int Nq, int Ns, int H, int K, float* __restrict__ wf,
                      float* __restrict__ inv_nn) {
    __shared__ float s_kp[KP_MAX_K * 3];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    if (threadIdx.x < K * 3) s_kp[threadIdx.x] = kp[threadIdx.x];
    __syncthreads();
    const int n = blockIdx.x * wpb + wib;
    if (n >= Nq) return;
    const float qx = q[3 * (size_t)n], qy = q[3 * (size_t)n + 1], qz = q[3 * (size_t)n + 2];
    const float ext2 = extent * extent, inv_ext = 1.0f / extent;
    float acc[KP_MAX_K];
#pragma unroll
    for (int k = 0; k < KP_MAX_K; ++k) acc[k] = 0.f;
    int nn = 0;
    for (int h = lane; h < H; h += 32) {
        const long long v = (long long)idx[(size_t)n * ld + h];
        if (v < 0 || v >= Ns) continue;
        const float4 p = __ldg(s4 + (int)v);           // (x, y, z, feature): with Cin == 1 the packed record carries x itself
        const float xv = p.w;
        nn += xv > 0.f ? 1 : 0;
        const float rx = p.x - qx, ry = p.y - qy, rz = p.z - qz;
#pragma unroll
        for (int k = 0; k < KP_MAX_K; ++k) {
            if (k < K) {
                const float ddx = rx - s_kp[3 * k], ddy = ry - s_kp[3 * k + 1], ddz = rz - s_kp[3 * k + 2];
                const float d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                if (d2 < ext2) acc[k] = fmaf(fmaxf(1.0f - sqrt_approx(d2) * inv_ext, 0.f), xv, acc[k]);
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
#pragma unroll
        for (int k = 0; k < KP_MAX_K; ++k) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], d);
    }
    nn = __reduce_add_sync(0xffffffffu, nn);
    if (lane == 0) {
        inv_nn[n] = 1.0f / (float)max(nn, 1);
#pragma unroll
        for (int k = 0; k < KP_MAX_K; ++k)
            if (k < K) wf[(size_t)n * K + k] = acc[k];
    }
}

// Plain fp32 tiled GEMM with row scale: C[M,N] = (A[M,K] @ B[K,N]) * rowscale[M]; A, B row-major. 64x64x16 tiles,
// 256 threads, 4x4 outputs per thread. The parity baseline of stage C (CUDA cores, exact fp32 products).
__global__ void __launch_bounds__(256)
sgemm_rowscale_kernel(const float* __restrict__ A, const float* __restrict__ B, int M, int N, int K,
                      const float* __restrict__ rowscale, float* __restrict__ C) {
    __shared__ float sA[16][64 + 4];
    __shared__ float sB[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int t = threadIdx.x; t < 64 * 16; t += 256) {
            int r = t >> 4, c = t & 15;  // A tile: 64 rows x 16 k
            int gm = m0 + r, gk = k0 + c;
            sA[c][r] = (gm < M && gk < K) ? A[(size_t)gm * K + gk] : 0.f;
            int kr = t >> 6, nc = t & 63;  // B tile: 16 k x 64 cols
            int gk2 = k0 + kr, gn = n0 + nc;
            sB[kr][nc] = (gk2 < K && gn < N) ? B[(size_t)gk2 * N + gn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
        float sc = rowscale ? rowscale[gm] : 1.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
            if (gn < N) C[(size_t)gm * N + gn] = acc[i][j] * sc;
        }
    }
}

// ---- training path: data gradient of stage A+B -------------------------------------------------
// dx[s,:] += sum over (n,h,k) with idx[n,h] == s of w[n,k,h] * dwf[n,k,:]   (the transpose of blocks.py:348-354).
// One warp per query: the same influence lists as kp_weighted_kernel (recomputed, nothing is stored by the forward),
// then each kernel point's dwf slice is held in registers and pushed to its list's support rows with red.add.f32.
template <typename IdxT, int VEC>
__global__ void __launch_bounds__(128)
kp_scatter_grad_kernel(const float* __restrict__ q, const float* __restrict__ s, const IdxT* __restrict__ idx, int ld,
                       const float* __restrict__ kp, float extent, int Nq, int Ns, int H, int K, int Cin,
                       const float* __restrict__ dwf, float* __restrict__ dx) {
    extern __shared__ float s_dyn[];
    __shared__ float s_kp[KP_MAX_K * 3];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int Hp = (H + 31) & ~31;
    if (threadIdx.x < K * 3) s_kp[threadIdx.x] = kp[threadIdx.x];
    __syncthreads();
    const int n = blockIdx.x * wpb + wib;
    if (n >= Nq) return;
    int* s_si = reinterpret_cast<int*>(s_dyn) + (size_t)wib * KP_MAX_K * Hp * 2;
    float* s_lw = reinterpret_cast<float*>(s_si + (size_t)KP_MAX_K * Hp);
    int* s_cnt = reinterpret_cast<int*>(s_dyn) + (size_t)wpb * KP_MAX_K * Hp * 2 + wib * KP_MAX_K;
    const float qx = q[3 * (size_t)n], qy = q[3 * (size_t)n + 1], qz = q[3 * (size_t)n + 2];
    const float ext2 = extent * extent, inv_ext = 1.0f / extent;
    int cnt[KP_MAX_K];
#pragma unroll
    for (int k = 0; k < KP_MAX_K; ++k) cnt[k] = 0;
    for (int h0 = 0; h0 < Hp; h0 += 32) {
        const int h = h0 + lane;
        int si = Ns;
        if (h < H) {
            const long long v = (long long)idx[(size_t)n * ld + h];
            si = (v >= 0 && v < Ns) ? (int)v : Ns;
        }
        float rx = 0.f, ry = 0.f, rz = 0.f;
        const bool valid = si < Ns;
        if (valid) { rx = s[3 * (size_t)si] - qx; ry = s[3 * (size_t)si + 1] - qy; rz = s[3 * (size_t)si + 2] - qz; }
        if (!__any_sync(0xffffffffu, valid)) continue;
#pragma unroll
        for (int k = 0; k < KP_MAX_K; ++k) {
            if (k < K) {
                const float ddx = rx - s_kp[3 * k], ddy = ry - s_kp[3 * k + 1], ddz = rz - s_kp[3 * k + 2];
                const float d2 = ddx * ddx + ddy * ddy + ddz * ddz;
                float w = 0.f;
                if (valid && d2 < ext2) w = 1.0f - sqrtf(d2) * inv_ext;
                const bool in = w > 0.f;
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (in) {
                    const int pos = cnt[k] + __popc(m & ((1u << lane) - 1));
                    s_si[k * Hp + pos] = si * Cin;
                    s_lw[k * Hp + pos] = w;
                }
                cnt[k] += __popc(m);
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < KP_MAX_K; ++k) s_cnt[k] = cnt[k];
    }
    __syncwarp();
    const float* grow = dwf + (size_t)n * K * Cin;
    for (int c0 = lane * VEC; c0 < Cin; c0 += 32 * VEC) {
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
            const int ck = s_cnt[k];
            if (ck == 0) continue;
            float g[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) g[v] = grow[(size_t)k * Cin + c0 + v];
            for (int e = 0; e < ck; ++e) {
                const float w = s_lw[k * Hp + e];
                float* dst = dx + s_si[k * Hp + e] + c0;
                if (VEC == 4) atomicAdd(reinterpret_cast<float4*>(dst), make_float4(w * g[0], w * g[VEC > 1 ? 1 : 0], w * g[VEC > 2 ? 2 : 0], w * g[VEC > 3 ? 3 : 0]));
                else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) atomicAdd(dst + v, w * g[v]);
                }
            }
        }
    }
}


// W [K,Cin,Cout] -> Wt [Cout, K*Cin] (K-major B operand), rounded to TF32 (round-to-nearest, ties away)
__global__ void prep_weights_kernel(const float* __restrict__ W, int KC, int Cout, float* __restrict__ Wt) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)KC * Cout) return;
    int o = (int)(t / KC), kc = (int)(t % KC);
    float v = W[(size_t)kc * Cout + o];
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    Wt[t] = __uint_as_float(u);
}

}  // namespace aprb

using namespace aprb;

// W [K,Cin,Cout] -> Wt [Cout, K*Cin] in fp16 (round to nearest even): B operand of the fp16-operand contraction (mode 3)
__global__ void prep_weights_f16_kernel(const float* __restrict__ W, int KC, int Cout, __half* __restrict__ Wt) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)KC * Cout) return;
    int o = (int)(t / KC), kc = (int)(t % KC);
    Wt[t] = __float2half_rn(W[(size_t)kc * Cout + o]);
}

// W [K,Cin,Cout] -> Wt [Cout, Cin*16] fp16 with column c*16 + k (zero for k >= K): B operand for the weighted tile of the
// tcgen05 weighting kernel (kpconv_tc.cu), which is channel-major / kernel-point-minor
__global__ void prep_weights_f16_ck_kernel(const float* __restrict__ W, int K, int Cin, int Cout, __half* __restrict__ Wt) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)Cin * KP_MAX_K * Cout) return;
    const int o = (int)(t / (Cin * KP_MAX_K)), ck = (int)(t % (Cin * KP_MAX_K));
    const int c = ck / KP_MAX_K, k = ck % KP_MAX_K;
    Wt[t] = k < K ? __float2half_rn(W[((size_t)k * Cin + c) * Cout + o]) : __float2half_rn(0.f);
}

extern "C" int aprb_kpconv_prepare_weights_f16_ck(const float* d_W, int K, int Cin, int Cout, void* d_wprep16ck, void* stream) {
    APRB_REQUIRE(d_W && d_wprep16ck && K >= 1 && K <= KP_MAX_K && Cin >= 1 && Cout >= 1, "bad argument");
    long long total = (long long)Cin * KP_MAX_K * Cout;
    APRB_TIMED("prep_weights_kernel", (cudaStream_t)stream, 1, (prep_weights_f16_ck_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(d_W, K, Cin, Cout, (__half*)d_wprep16ck)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

extern "C" int aprb_kpconv_tc_supported(int H, int K, int Cin, int Cout, int Ns) {
    return kpconv_tc_supported(H, K, Cin, Ns) && gemm_f16_supported(1, Cout, KP_MAX_K * Cin) ? 1 : 0;
}

extern "C" int aprb_kpconv_prepare_weights_f16(const float* d_W, int K, int Cin, int Cout, void* d_wprep16, void* stream) {
    APRB_REQUIRE(d_W && d_wprep16 && K >= 1 && Cin >= 1 && Cout >= 1, "bad argument");
    long long total = (long long)K * Cin * Cout;
    APRB_TIMED("prep_weights_kernel", (cudaStream_t)stream, 1, (prep_weights_f16_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(d_W, K * Cin, Cout, (__half*)d_wprep16)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

extern "C" int aprb_kpconv_prepare_weights(const float* d_W, int K, int Cin, int Cout, float* d_wprep, void* stream) {
    APRB_REQUIRE(d_W && d_wprep && K >= 1 && Cin >= 1 && Cout >= 1, "bad argument");
    long long total = (long long)K * Cin * Cout;
    APRB_TIMED("prep_weights_kernel", (cudaStream_t)stream, 1, (prep_weights_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(d_W, K * Cin, Cout, d_wprep)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

__global__ void round_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(in[i]));
    out[i] = __uint_as_float(u);
}

extern "C" int aprb_round_tf32(const float* d_in, float* d_out, size_t n, void* stream) {
    APRB_REQUIRE(n == 0 || (d_in && d_out), "null pointer");
    if (n == 0) return APRB_OK;
    APRB_TIMED("round_tf32_kernel", (cudaStream_t)stream, 1, (round_tf32_kernel<<<cdiv((long long)n, 256), 256, 0, (cudaStream_t)stream>>>(d_in, d_out, n)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

// Launch stage A+B for query rows [r0, r0 + nr): wf rows are written relative to r0, inv_nn at absolute rows.
static int launch_kp_weighted(const float* d_q, const float* d_s, const void* d_idx, int idx_is_i64, int ld_idx,
                              const float* d_x, const float* d_kp, const unsigned char* flag, const float4* s4, float extent, int r0, int nr,
                              int Ns, int H, int K, int Cin, bool round_tf32, float* wf, float* inv_nn, cudaStream_t st, int out16 = 0, int x16 = 0) {
    const int Hp = (H + 31) & ~31;
    const bool al16 = ((uintptr_t)d_x % 16 == 0) && ((uintptr_t)wf % 16 == 0);
    if (out16 && !(al16 && Cin % 4 == 0 && H <= 128 && (long long)Ns * Cin < (1LL << 30))) {
        set_error("aprb_kpconv_forward: fp16 operand path needs Cin %% 4 == 0, H <= 128 and 16-byte aligned features");
        return APRB_ERR_UNSUPPORTED;
    }
    if ((g_kpw_version >= 4 || out16) && al16 && Cin % 4 == 0 && H <= 128 && (long long)Ns * Cin < (1LL << 30)) {
        // v4: lane-group streaming over compact CSR lists; (LG, NV) by channel count, NH = 32-neighbour groups per row
        const int nh = H <= 64 ? 2 : 4;
        const int chs[5] = {32, 64, 128, 256, 512};
        int cfg = Cin <= 32 ? 0 : (Cin <= 64 ? 1 : (Cin <= 128 ? 2 : (Cin <= 256 ? 3 : 4)));
        const bool full = Cin % chs[cfg] == 0;
#define KPW4_LAUNCH(IDX, LG, NV, NH, FULLV, RND, XH)                                                                             \
        do {                                                                                                          \
            constexpr int wpb4 = 4;                                                                                   \
            constexpr int rp = 32 / LG;                                                                               \
            const size_t smem4 = (size_t)wpb4 * rp * KPW_SLOT_BYTES;                                                  \
            APRB_TIMED("kp_weighted_kernel", st, 1, (kp_weighted4_kernel<IDX, LG, NV, NH, FULLV, RND, XH><<<cdiv(nr, wpb4 * rp), wpb4 * 32, smem4, st>>>( \
                d_q + 3 * (size_t)r0, s4, (const IDX*)d_idx + (size_t)r0 * ld_idx, ld_idx, d_x, d_kp, extent, nr, Ns, H, K, Cin, wf, inv_nn + r0, out16))); \
        } while (0)
#define KPW4_CFG(IDX, NH, RND)                                                                                        \
        do {                                                                                                          \
            if (cfg == 0) { if (full) KPW4_LAUNCH(IDX, 8, 1, NH, true, RND, false); else KPW4_LAUNCH(IDX, 8, 1, NH, false, RND, false); }                                                            \
            else if (cfg == 1) { if (full) KPW4_LAUNCH(IDX, 16, 1, NH, true, RND, false); else KPW4_LAUNCH(IDX, 16, 1, NH, false, RND, false); }                                                      \
            else if (cfg == 2) { if (full) KPW4_LAUNCH(IDX, 16, 2, NH, true, RND, false); else KPW4_LAUNCH(IDX, 16, 2, NH, false, RND, false); }                                                      \
            else if (cfg == 3) { if (full) KPW4_LAUNCH(IDX, 32, 2, NH, true, RND, false); else KPW4_LAUNCH(IDX, 32, 2, NH, false, RND, false); }                                                      \
            else { if (full) KPW4_LAUNCH(IDX, 32, 4, NH, true, RND, false); else KPW4_LAUNCH(IDX, 32, 4, NH, false, RND, false); }                                                                    \
        } while (0)
#define KPW4_NH(IDX, RND) do { if (nh == 2) KPW4_CFG(IDX, 2, RND); else KPW4_CFG(IDX, 4, RND); } while (0)
        if (x16) {
            // fp16 features in, fp16 weighted tile out (the native pipeline's activation format): int32 indices, whole slabs
            if (idx_is_i64 || !full || !out16) { set_error("aprb_kpconv_forward: fp16 features need int32 indices and Cin %% %d == 0", chs[cfg]); return APRB_ERR_UNSUPPORTED; }
            if (g_kpw_fh && Cin % 64 == 0 && Ns > 0) {
                // v5 (FHFMA): 8 channels per lane and load; Cin = 64 runs 4 rows per warp (g_kpw_fh == 2: 2 rows, 4 channels per lane)
#define KPWH_LAUNCH(LG, CPL, NV, NH)                                                                                  \
                do {                                                                                                  \
                    constexpr int wpbh = 4;                                                                           \
                    constexpr int rph = 32 / LG;                                                                      \
                    const size_t smemh = (size_t)wpbh * rph * KPW_SLOT_BYTES;                                         \
                    APRB_TIMED("kp_weighted_kernel", st, 1, (kp_weighted_h_kernel<LG, CPL, NV, NH><<<cdiv(nr, wpbh * rph), wpbh * 32, smemh, st>>>( \
                        d_q + 3 * (size_t)r0, s4, (const int*)d_idx + (size_t)r0 * ld_idx, ld_idx, (const __half*)d_x, d_kp, extent, nr, Ns, H, K, Cin, \
                        (__half*)wf, inv_nn + r0)));                                                                  \
                } while (0)
#define KPWH_CFG(NH)                                                                                                  \
                do {                                                                                                  \
                    if (Cin == 64) { if (g_kpw_fh == 2) KPWH_LAUNCH(16, 4, 1, NH); else KPWH_LAUNCH(8, 8, 1, NH); }   \
                    else if (Cin == 128) KPWH_LAUNCH(16, 8, 1, NH);                                                   \
                    else if (Cin % 512 == 0) KPWH_LAUNCH(32, 8, 2, NH);                                               \
                    else if (Cin % 256 == 0) KPWH_LAUNCH(32, 8, 1, NH);                                               \
                    else KPWH_LAUNCH(8, 8, 1, NH);                                                                    \
                } while (0)
                if (nh == 2) KPWH_CFG(2); else KPWH_CFG(4);
#undef KPWH_CFG
#undef KPWH_LAUNCH
                APRB_LAUNCH_OK();
                return APRB_OK;
            }
#define KPW4_X16(NH)                                                                                                  \
            do {                                                                                                      \
                if (cfg == 0) KPW4_LAUNCH(int, 8, 1, NH, true, false, true);                                          \
                else if (cfg == 1) KPW4_LAUNCH(int, 16, 1, NH, true, false, true);                                    \
                else if (cfg == 2) KPW4_LAUNCH(int, 16, 2, NH, true, false, true);                                    \
                else if (cfg == 3) KPW4_LAUNCH(int, 32, 2, NH, true, false, true);                                    \
                else KPW4_LAUNCH(int, 32, 4, NH, true, false, true);                                                  \
            } while (0)
            if (nh == 2) KPW4_X16(2); else KPW4_X16(4);
#undef KPW4_X16
        } else if (round_tf32) { if (idx_is_i64) KPW4_NH(long long, true); else KPW4_NH(int, true); }
        else { if (idx_is_i64) KPW4_NH(long long, false); else KPW4_NH(int, false); }
#undef KPW4_NH
#undef KPW4_CFG
#undef KPW4_LAUNCH
        APRB_LAUNCH_OK();
        return APRB_OK;
    }
    const size_t smem_warp = (size_t)KP_MAX_K * Hp * 8 + KP_MAX_K * 4;
    int wpb = 4;
    while (wpb > 1 && wpb * smem_warp > 160 * 1024) wpb >>= 1;
    if (smem_warp > 200 * 1024) { set_error("aprb_kpconv_forward: H=%d too large for the shared-memory neighbour lists", H); return APRB_ERR_UNSUPPORTED; }
    const size_t smem = wpb * smem_warp;
    // (VEC, NJ): channels per lane = VEC*NJ, one slab = 32*VEC*NJ channels
    int vec = 1, nj = 1;
    if (al16 && Cin % 4 == 0 && Cin >= 128) { vec = 4; nj = Cin >= 512 ? 4 : (Cin >= 256 ? 2 : 1); }
    else if (al16 && Cin % 2 == 0 && Cin >= 64) { vec = 2; nj = 1; }
#define KPW_LAUNCH3(IDX, VEC, NJ, RND)                                                                               \
    do {                                                                                                             \
        if (smem > 48 * 1024)                                                                                        \
            APRB_CUDA_OK(cudaFuncSetAttribute(kp_weighted_kernel<IDX, VEC, NJ, RND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        const int nslabs = cdiv(Cin, 32 * VEC * NJ);                                                                 \
        const int gy = (cdiv(nr, wpb) < 4 * sm_count()) ? nslabs : 1;                                                \
        APRB_TIMED("kp_weighted_kernel", st, 1, (kp_weighted_kernel<IDX, VEC, NJ, RND><<<dim3(cdiv(nr, wpb), gy), wpb * 32, smem, st>>>( \
            d_q + 3 * (size_t)r0, d_s, (const IDX*)d_idx + (size_t)r0 * ld_idx, ld_idx, d_x, d_kp, flag, extent, nr, Ns, H, K, Cin, wf, inv_nn + r0))); \
    } while (0)
#define KPW_LAUNCH(IDX, RND)                                                                                         \
    do {                                                                                                             \
        if (vec == 4 && nj == 4) KPW_LAUNCH3(IDX, 4, 4, RND);                                                        \
        else if (vec == 4 && nj == 2) KPW_LAUNCH3(IDX, 4, 2, RND);                                                   \
        else if (vec == 4) KPW_LAUNCH3(IDX, 4, 1, RND);                                                              \
        else if (vec == 2) KPW_LAUNCH3(IDX, 2, 1, RND);                                                              \
        else KPW_LAUNCH3(IDX, 1, 1, RND);                                                                            \
    } while (0)
    if (round_tf32) { if (idx_is_i64) KPW_LAUNCH(long long, true); else KPW_LAUNCH(int, true); }
    else { if (idx_is_i64) KPW_LAUNCH(long long, false); else KPW_LAUNCH(int, false); }
#undef KPW_LAUNCH3
#undef KPW_LAUNCH
    APRB_LAUNCH_OK();
    return APRB_OK;
}

extern "C" size_t aprb_kpconv_ws_bytes(int Nq, int Ns, int H, int K, int Cin, int Cout) {
    (void)H;
    if (Nq < 0 || Ns < 0 || K < 0 || Cin < 0) return 0;
    // wf is padded to a multiple of 128 rows so the tensor path can read whole M tiles
    size_t rows = ((size_t)(Nq > 0 ? Nq : 1) + 127) & ~size_t(127);
    return align256(rows * K * Cin * sizeof(float)) + align256(rows * sizeof(float)) + align256((size_t)Ns + 1) +
           align256(((size_t)Ns + 1) * sizeof(float4)) +
           gemm_tf32_ws_bytes(Nq > 0 ? Nq : 1, Cout > 0 ? Cout : 1) + 256;
}

extern "C" int aprb_kpconv_forward(const float* d_q, const float* d_s, const void* d_idx, int idx_is_i64, int ld_idx,
                                   const float* d_x, const float* d_kp, const float* d_W, const float* d_wprep,
                                   float extent, int Nq, int Ns, int H, int K, int Cin, int Cout, float* d_out, int mode,
                                   void* d_ws, size_t ws_bytes, void* stream) {
    return aprb_kpconv_forward_stats(d_q, d_s, d_idx, idx_is_i64, ld_idx, d_x, d_kp, d_W, d_wprep, extent, Nq, Ns, H, K, Cin,
                                     Cout, d_out, mode, nullptr, nullptr, d_ws, ws_bytes, stream);
}

extern "C" int aprb_kpconv_forward_stats(const float* d_q, const float* d_s, const void* d_idx, int idx_is_i64, int ld_idx,
                                         const float* d_x, const float* d_kp, const float* d_W, const float* d_wprep,
                                         float extent, int Nq, int Ns, int H, int K, int Cin, int Cout, float* d_out, int mode,
                                         float* d_gstat, int* stats_written, void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (stats_written) *stats_written = 0;
    APRB_REQUIRE(!d_gstat || stats_written, "stats_written must be given with d_gstat");
    APRB_REQUIRE(Nq >= 0 && Ns >= 0 && H >= 1 && H <= 1024, "need Nq,Ns >= 0 and 1 <= H <= 1024");
    APRB_REQUIRE(K >= 1 && K <= KP_MAX_K && Cin >= 1 && Cout >= 1 && ld_idx >= H, "need 1 <= K <= 16, Cin,Cout >= 1, ld >= H");
    APRB_REQUIRE(extent > 0.f && extent < 1e15f, "extent must be positive (and below 1e15)");
    APRB_REQUIRE((long long)Ns * Cin < 0x7FFFFFFFLL, "feature table too large for 32-bit row offsets");
    if (Nq == 0) return APRB_OK;
    APRB_REQUIRE(d_q && d_idx && d_kp && d_out && d_ws && (Ns == 0 || (d_s && d_x)), "null pointer");
    APRB_REQUIRE(mode >= 0 && mode <= 5, "mode must be 0 .. 5");
    if (ws_bytes < aprb_kpconv_ws_bytes(Nq, Ns, H, K, Cin, Cout)) { set_error("aprb_kpconv_forward: workspace too small"); return APRB_ERR_WORKSPACE; }
    const int KC = K * Cin;
    if (mode == 5) {
        // fp16 features, weighting stage on tcgen05 (kpconv_tc.cu), weighted tile [Nq, Cin*16] channel-major, contraction
        // against the ck-ordered fp16 operand (aprb_kpconv_prepare_weights_f16_ck)
        APRB_REQUIRE(d_wprep && !idx_is_i64, "mode 5 needs the ck-ordered fp16 prepared weights and int32 indices");
        if (Ns == 0 || !kpconv_tc_supported(H, K, Cin, Ns) || !gemm_f16_supported(Nq, Cout, KP_MAX_K * Cin)) {
            set_error("aprb_kpconv_forward: mode 5 unsupported for H=%d Cin=%d Cout=%d", H, Cin, Cout);
            return APRB_ERR_UNSUPPORTED;
        }
        Carver c5(d_ws, ws_bytes);
        size_t rows5 = ((size_t)Nq + 127) & ~size_t(127);
        float* wf5 = c5.take<float>(rows5 * KC);                      // same carving as the other modes; holds Nq*Cin*16 halves
        float* inv5 = c5.take<float>(rows5);
        unsigned char* flag5 = c5.take<unsigned char>((size_t)Ns + 1);
        float4* s45 = c5.take<float4>((size_t)Ns + 1);
        APRB_REQUIRE((((uintptr_t)d_x | (uintptr_t)wf5) & 15) == 0, "mode 5 needs 16-byte aligned features");
        launch_rowsum_pos(d_x, d_s, Ns, Cin, flag5, s45, 1, st);
        int rc = kpconv_tc_run(d_q, s45, (const int*)d_idx, ld_idx, d_x, d_kp, extent, Nq, Ns, H, K, Cin, wf5, inv5, st);
        if (rc) return rc;
        set_gemm_label("kpconv_gemm_kernel");
        rc = gemm_f16_rowscale(wf5, d_wprep, Nq, Cout, KP_MAX_K * Cin, inv5, d_out, st, d_gstat, stats_written);
        set_gemm_label(nullptr);
        return rc;
    }
    if (mode == 3 || mode == 4) {
        const int x16 = mode == 4;                                    // mode 4: d_x itself is fp16 [Ns, Cin]
        // tcgen05 with fp16 operands: d_wprep is the fp16 prepared operand (aprb_kpconv_prepare_weights_f16); the weighted
        // tile is produced in fp16 (same 10-bit mantissa as TF32, half the bytes of the largest tensor of the path).
        // fp16 range: the weighted tile saturates at +-65504 (pack_half2_sat) instead of overflowing to inf.
        APRB_REQUIRE(d_wprep, "mode 3 needs the fp16 prepared weights");
        if (!gemm_f16_supported(Nq, Cout, KC)) { set_error("aprb_kpconv_forward: fp16 path unsupported for K*Cin=%d Cout=%d", KC, Cout); return APRB_ERR_UNSUPPORTED; }
        Carver c3(d_ws, ws_bytes);
        size_t rows3 = ((size_t)Nq + 127) & ~size_t(127);
        float* wf3 = c3.take<float>(rows3 * KC);                      // same carving as below; holds Nq*KC halves
        float* inv3 = c3.take<float>(rows3);
        unsigned char* flag3 = c3.take<unsigned char>((size_t)Ns + 1);
        float4* s43 = c3.take<float4>((size_t)Ns + 1);
        if (Ns > 0) launch_rowsum_pos(d_x, d_s, Ns, Cin, flag3, s43, x16, st);
        int rc = launch_kp_weighted(d_q, d_s, d_idx, idx_is_i64, ld_idx, d_x, d_kp, flag3, s43, extent, 0, Nq, Ns, H, K, Cin, false, wf3, inv3, st, 1, x16);
        if (rc) return rc;
        set_gemm_label("kpconv_gemm_kernel");
        rc = gemm_f16_rowscale(wf3, d_wprep, Nq, Cout, KC, inv3, d_out, st, d_gstat, stats_written);
        set_gemm_label(nullptr);
        return rc;
    }
    bool tensor_ok = d_wprep && gemm_tf32_supported(Nq, Cout, KC);
    if (mode == 2 && !tensor_ok) { set_error("aprb_kpconv_forward: tcgen05 path unsupported for K*Cin=%d Cout=%d", KC, Cout); return APRB_ERR_UNSUPPORTED; }
    bool use_tensor = (mode == 2) || (mode == 0 && tensor_ok);
    if (!use_tensor) APRB_REQUIRE(d_W, "fp32 path needs the raw [K,Cin,Cout] weights");

    Carver c(d_ws, ws_bytes);
    size_t rows = ((size_t)Nq + 127) & ~size_t(127);
    float* wf = c.take<float>(rows * KC);
    float* inv_nn = c.take<float>(rows);
    unsigned char* flag = c.take<unsigned char>((size_t)Ns + 1);
    float4* s4 = c.take<float4>((size_t)Ns + 1);
    const size_t gws_bytes = gemm_tf32_ws_bytes(Nq, Cout) - 256;
    float* gws = c.take<float>(gws_bytes / sizeof(float));

    if (Ns > 0) APRB_TIMED("rowsum_pos_kernel", st, 1, (rowsum_pos_kernel<<<cdiv(Ns, 8), 256, 0, st>>>(d_x, d_s, Ns, Cin, flag, s4, 0)));
    if (use_tensor && Ns > 0 && kpconv_fused_supported(H, K, Cin, Cout, Ns) &&
        ((((uintptr_t)d_x | (uintptr_t)d_out | (uintptr_t)d_wprep | (uintptr_t)(d_gstat ? d_gstat : d_out)) & 15) == 0)) {
        // one kernel: gather -> influence -> swizzled A tiles in shared memory -> tcgen05 (kpconv_fused.cu)
        if (d_gstat) *stats_written = 1;
        return kpconv_fused_run(d_q, s4, d_idx, idx_is_i64, ld_idx, d_x, d_kp, d_wprep, extent, Nq, Ns, H, K, Cin, Cout, d_out,
                                d_gstat, st);
    }
    if (use_tensor) {
        // Row chunks sized so that one chunk of wf (the A operand) stays L2-resident between its producer and the GEMM
        // that consumes it (aprb_set_option("kpconv_chunk_mb"); measured on B200: no gain, off by default).
        int chunk_rows = Nq;
        if (g_kpconv_chunk_mb > 0) {
            long long rows_c = ((long long)g_kpconv_chunk_mb << 20) / ((long long)KC * 4);
            rows_c = (rows_c / 128) * 128;
            if (rows_c < 128 * 148) rows_c = 128 * 148;              // at least one GEMM tile per SM
            if (rows_c < Nq) chunk_rows = (int)rows_c;
        }
        for (int r0 = 0; r0 < Nq; r0 += chunk_rows) {
            const int nr = min(chunk_rows, Nq - r0);
            int rc = launch_kp_weighted(d_q, d_s, d_idx, idx_is_i64, ld_idx, d_x, d_kp, flag, s4, extent, r0, nr, Ns, H, K, Cin, true, wf, inv_nn, st);
            if (rc) return rc;
            // group statistics only when the operator is one GEMM over all rows (row chunks would misalign the groups)
            set_gemm_label("kpconv_gemm_kernel");
            rc = gemm_tf32_rowscale(wf, d_wprep, nr, Cout, KC, inv_nn + r0, d_out + (size_t)r0 * Cout, gws, gws_bytes, st,
                                    chunk_rows == Nq ? d_gstat : nullptr, chunk_rows == Nq ? stats_written : nullptr);
            set_gemm_label(nullptr);
            if (rc) return rc;
        }
        return APRB_OK;
    }
    if (Cin == 1) {
        if (idx_is_i64) APRB_TIMED("kp_weighted_c1_kernel", st, 1, (kp_weighted_c1_kernel<long long><<<cdiv(Nq, 4), 128, 0, st>>>(
            d_q, s4, (const long long*)d_idx, ld_idx, d_kp, extent, Nq, Ns, H, K, wf, inv_nn)));
        else APRB_TIMED("kp_weighted_c1_kernel", st, 1, (kp_weighted_c1_kernel<int><<<cdiv(Nq, 4), 128, 0, st>>>(
            d_q, s4, (const int*)d_idx, ld_idx, d_kp, extent, Nq, Ns, H, K, wf, inv_nn)));
        APRB_LAUNCH_OK();
    } else {
        int rc = launch_kp_weighted(d_q, d_s, d_idx, idx_is_i64, ld_idx, d_x, d_kp, flag, s4, extent, 0, Nq, Ns, H, K, Cin, false, wf, inv_nn, st);
        if (rc) return rc;
    }
    APRB_TIMED("sgemm_rowscale_kernel", st, 1, (sgemm_rowscale_kernel<<<dim3(cdiv(Cout, 64), cdiv(Nq, 64)), 256, 0, st>>>(wf, d_W, Nq, Cout, KC, inv_nn, d_out)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

extern "C" size_t aprb_kpconv_weighted_ws_bytes(int Ns) {
    return Ns < 0 ? 0 : align256((size_t)Ns + 1) + align256(((size_t)Ns + 1) * sizeof(float4)) + 256;
}

extern "C" int aprb_kpconv_weighted(const float* d_q, const float* d_s, const void* d_idx, int idx_is_i64, int ld_idx,
                                    const float* d_x, const float* d_kp, float extent, int Nq, int Ns, int H, int K, int Cin,
                                    int round_tf32, float* d_wf, float* d_inv_nn, void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(Nq >= 0 && Ns >= 0 && H >= 1 && H <= 1024, "need Nq,Ns >= 0 and 1 <= H <= 1024");
    APRB_REQUIRE(K >= 1 && K <= KP_MAX_K && Cin >= 1 && ld_idx >= H && extent > 0.f, "need 1 <= K <= 16, Cin >= 1, ld >= H, extent > 0");
    APRB_REQUIRE((long long)Ns * Cin < 0x7FFFFFFFLL, "feature table too large for 32-bit row offsets");
    if (Nq == 0) return APRB_OK;
    APRB_REQUIRE(d_q && d_idx && d_kp && d_wf && d_inv_nn && d_ws && (Ns == 0 || (d_s && d_x)), "null pointer");
    if (ws_bytes < aprb_kpconv_weighted_ws_bytes(Ns)) { set_error("aprb_kpconv_weighted: workspace too small"); return APRB_ERR_WORKSPACE; }
    unsigned char* flag = (unsigned char*)d_ws;
    float4* s4 = (float4*)((char*)d_ws + align256((size_t)Ns + 1));
    if (Ns > 0) APRB_TIMED("rowsum_pos_kernel", st, 1, (rowsum_pos_kernel<<<cdiv(Ns, 8), 256, 0, st>>>(d_x, d_s, Ns, Cin, flag, s4, 0)));
    if (Cin == 1) {
        if (idx_is_i64) APRB_TIMED("kp_weighted_c1_kernel", st, 1, (kp_weighted_c1_kernel<long long><<<cdiv(Nq, 4), 128, 0, st>>>(
            d_q, s4, (const long long*)d_idx, ld_idx, d_kp, extent, Nq, Ns, H, K, d_wf, d_inv_nn)));
        else APRB_TIMED("kp_weighted_c1_kernel", st, 1, (kp_weighted_c1_kernel<int><<<cdiv(Nq, 4), 128, 0, st>>>(
            d_q, s4, (const int*)d_idx, ld_idx, d_kp, extent, Nq, Ns, H, K, d_wf, d_inv_nn)));
        APRB_LAUNCH_OK();
        return APRB_OK;
    }
    return launch_kp_weighted(d_q, d_s, d_idx, idx_is_i64, ld_idx, d_x, d_kp, flag, s4, extent, 0, Nq, Ns, H, K, Cin, round_tf32 != 0,
                              d_wf, d_inv_nn, st);
}

// Stage A+B alone in the native pipeline's format: fp16 features in, fp16 weighted tile out. layout_ck = 0: [Nq, K*Cin]
// kernel-point-major by the CUDA-core list kernel (mode 4 of aprb_kpconv_forward); 1: [Nq, Cin*16] channel-major by the
// tcgen05 weighting kernel (mode 5).
extern "C" int aprb_kpconv_weighted_f16(const float* d_q, const float* d_s, const int32_t* d_idx, int ld_idx, const void* d_x16,
                                        const float* d_kp, float extent, int Nq, int Ns, int H, int K, int Cin, int layout_ck,
                                        void* d_wf16, float* d_inv_nn, void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(Nq >= 0 && Ns >= 0 && H >= 1 && H <= 128, "need Nq,Ns >= 0 and 1 <= H <= 128");
    APRB_REQUIRE(K >= 1 && K <= KP_MAX_K && Cin >= 4 && Cin % 4 == 0 && ld_idx >= H && extent > 0.f, "need 1 <= K <= 16, Cin % 4 == 0, ld >= H, extent > 0");
    APRB_REQUIRE((long long)Ns * Cin < (1LL << 30), "feature table too large for 32-bit row offsets");
    if (Nq == 0) return APRB_OK;
    APRB_REQUIRE(d_q && d_idx && d_kp && d_wf16 && d_inv_nn && d_ws && (Ns == 0 || (d_s && d_x16)), "null pointer");
    if (ws_bytes < aprb_kpconv_weighted_ws_bytes(Ns)) { set_error("aprb_kpconv_weighted_f16: workspace too small"); return APRB_ERR_WORKSPACE; }
    unsigned char* flag = (unsigned char*)d_ws;
    float4* s4 = (float4*)((char*)d_ws + align256((size_t)Ns + 1));
    if (Ns > 0) launch_rowsum_pos((const float*)d_x16, d_s, Ns, Cin, flag, s4, 1, st);
    if (layout_ck) {
        if (!(Ns > 0 && kpconv_tc_supported(H, K, Cin, Ns) && (((uintptr_t)d_x16 | (uintptr_t)d_wf16) & 15) == 0)) {
            set_error("aprb_kpconv_weighted_f16: the tcgen05 weighting kernel does not support H=%d Cin=%d", H, Cin);
            return APRB_ERR_UNSUPPORTED;
        }
        return kpconv_tc_run(d_q, s4, d_idx, ld_idx, d_x16, d_kp, extent, Nq, Ns, H, K, Cin, d_wf16, d_inv_nn, st);
    }
    return launch_kp_weighted(d_q, d_s, d_idx, 0, ld_idx, (const float*)d_x16, d_kp, flag, s4, extent, 0, Nq, Ns, H, K, Cin, false,
                              (float*)d_wf16, d_inv_nn, st, 1, 1);
}

extern "C" int aprb_kpconv_backward_data(const float* d_q, const float* d_s, const void* d_idx, int idx_is_i64, int ld_idx,
                                         const float* d_kp, float extent, int Nq, int Ns, int H, int K, int Cin,
                                         const float* d_dwf, float* d_dx, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(Nq >= 0 && Ns >= 0 && H >= 1 && H <= 1024, "need Nq,Ns >= 0 and 1 <= H <= 1024");
    APRB_REQUIRE(K >= 1 && K <= KP_MAX_K && Cin >= 1 && ld_idx >= H && extent > 0.f, "need 1 <= K <= 16, Cin >= 1, ld >= H, extent > 0");
    APRB_REQUIRE((long long)Ns * Cin < 0x7FFFFFFFLL, "feature table too large for 32-bit row offsets");
    if (Ns == 0) return APRB_OK;
    APRB_REQUIRE(d_dx, "null pointer");
    APRB_CUDA_OK(cudaMemsetAsync(d_dx, 0, (size_t)Ns * Cin * sizeof(float), st));
    if (Nq == 0) return APRB_OK;
    APRB_REQUIRE(d_q && d_s && d_idx && d_kp && d_dwf, "null pointer");
    const int Hp = (H + 31) & ~31;
    const size_t smem_warp = (size_t)KP_MAX_K * Hp * 8 + KP_MAX_K * 4;
    int wpb = 4;
    while (wpb > 1 && wpb * smem_warp > 160 * 1024) wpb >>= 1;
    if (smem_warp > 200 * 1024) { set_error("aprb_kpconv_backward_data: H=%d too large", H); return APRB_ERR_UNSUPPORTED; }
    const size_t smem = wpb * smem_warp;
    const bool v4 = Cin % 4 == 0 && ((uintptr_t)d_dwf % 16 == 0) && ((uintptr_t)d_dx % 16 == 0);
#define KPS_LAUNCH(IDX, VEC)                                                                                          \
    do {                                                                                                              \
        if (smem > 48 * 1024)                                                                                         \
            APRB_CUDA_OK(cudaFuncSetAttribute(kp_scatter_grad_kernel<IDX, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        APRB_TIMED("kp_scatter_grad_kernel", st, 1, (kp_scatter_grad_kernel<IDX, VEC><<<cdiv(Nq, wpb), wpb * 32, smem, st>>>( \
            d_q, d_s, (const IDX*)d_idx, ld_idx, d_kp, extent, Nq, Ns, H, K, Cin, d_dwf, d_dx)));                      \
    } while (0)
    if (idx_is_i64) { if (v4) KPS_LAUNCH(long long, 4); else KPS_LAUNCH(long long, 1); }
    else { if (v4) KPS_LAUNCH(int, 4); else KPS_LAUNCH(int, 1); }
#undef KPS_LAUNCH
    APRB_LAUNCH_OK();
    return APRB_OK;
}
