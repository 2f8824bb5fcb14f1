// kpconv_fused.cu — K5 fused: KPConv.forward (/root/reference/Predator_APR/models/blocks.py:229-374, rigid kernel points,
// linear influence, sum aggregation) as ONE kernel per 128-query tile: neighbour gather -> kernel-point influence ->
// [128 x 32] slices of the weighted tile written straight into 128B-swizzled shared memory -> tcgen05.mma (TF32, fp32
// accumulation in TMEM) against TMA-streamed slices of the prepared weights [Cout, K*Cin] -> epilogue with the
// 1/neighbor_num row scale and the group statistics of the InstanceNorm that follows. The [Nq, K*Cin] intermediate
// ("wf", 3.8 - 15 KB per query, 11.6 GB of HBM traffic per 8-pair KFE forward when written and re-read) never exists.
//
// CTA = 16 producer warps + 1 TMA warp + 1 MMA warp, one 128-row tile:
//   phase 1  each producer warp builds the CSR influence lists of its 8 rows in shared memory (kpconv_common.cuh,
//            the same code as the stand-alone producer kp_weighted4_kernel); lists are warp-private, so no CTA barrier;
//   phase 2  kernel-point-major passes p = (k, 64-channel slab): two lane groups of a warp stream two rows at a time
//            (128-bit gathers, fp32 FMA), pre-round to TF32 and store the row slices into two A stages (one per
//            32-channel k-block) in the layout TMA + SWIZZLE_128B would have produced (16-byte chunk j of row r at
//            chunk j ^ (r & 7)); fence.proxy.async, then one mbarrier arrive per warp and stage;
//   TMA warp streams the matching [Cout x 32] weight k-blocks; the MMA warp waits for both, issues 4 x (M128, N=Cout,
//            K8) per k-block into TMEM and commits the stages back to the producers / the TMA warp;
//   epilogue warps 0-3 (their rows are done first) drain TMEM through a transpose buffer that aliases the A ring.
// A row whose list exceeds KPF_ECAP entries (kernel points much closer together than the extent) keeps no list: the
// whole warp re-evaluates its influences in every pass (direct path), so any kernel-point configuration is exact.
#include "kpconv_common.cuh"
#include "tc_common.cuh"

namespace aprb {

// Row slot: byte offsets of the row's (up to 64) neighbours' feature rows, then the CSR entries packed in 4 bytes —
// the influence weight with its 6 low mantissa bits replaced by the neighbour's position in the row (w keeps 17
// mantissa bits, 2^-18 relative, far below the TF32 rounding that follows) — then the K+1 list offsets.
constexpr int KPF_ECAP = 144;                      // entries per row (rows with all 56 neighbours average ~78)
constexpr int KPF_LROW = 256 + KPF_ECAP * 4 + 80;  // uint otab[64]; uint ent[ECAP]; int off[KP_MAX_K + 1] (+pad) = 912 bytes
constexpr int KPF_NPW = 16;                        // producer warps: 8 rows each
constexpr int KPF_NSA = 4;                         // A stages (two passes in flight)
constexpr int KPF_THREADS = (KPF_NPW + 2) * 32;

template <int COUT>
struct KpfCfg {
    static constexpr int NSB = COUT <= 64 ? 4 : 2;                 // B stages
    static constexpr int A_BYTES = GEMM_BM * 128;
    static constexpr int B_BYTES = COUT * 128;
    static constexpr int LISTS = GEMM_BM * KPF_LROW;
    static constexpr int SMEM = KPF_NSA * A_BYTES + NSB * B_BYTES + LISTS + 1024 /*align*/ + 512 /*inv_nn*/ + 256 /*barriers*/;
};

// build_row_list (kpconv_common.cuh) with the packed entry format above
template <int NH>
__device__ __forceinline__ void build_row_list_packed(const RowGeom<NH>& g, const float4* s_kp, int K, float ext2, float inv_ext,
                                                      uint32_t* otab, uint32_t* ent, int* off, int lane) {
    const unsigned ltmask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < 2; ++j) otab[j * 32 + lane] = j < NH ? (uint32_t)g.sio[j] : 0u;
    int run = 0, myoff = 0;
#pragma unroll
    for (int k = 0; k < KP_MAX_K; ++k) {
        if (lane == k) myoff = run;
        if (k < K) {
            const float4 kpk = s_kp[k];
#pragma unroll
            for (int j = 0; j < NH; ++j) {
                if (g.any[j]) {
                    bool in;
                    const float w = influence<NH>(g, j, kpk, ext2, inv_ext, in);
                    const unsigned m = __ballot_sync(0xffffffffu, in);
                    const int pos = run + __popc(m & ltmask);
                    if (in && pos < KPF_ECAP) ent[pos] = ((__float_as_uint(w) + 32u) & ~63u) | (uint32_t)(j * 32 + lane);
                    run += __popc(m);
                }
            }
        }
    }
    if (lane >= K) myoff = run;
    if (lane <= KP_MAX_K) off[lane] = myoff;
}

// Direct (list-free) evaluation of one row's slice of pass (k, slab) by the whole warp: lanes = neighbours for the
// influence, lanes 0-15 = the 16 float4 of the 64-channel slab for the accumulation.
template <typename IdxT, int NH>
__device__ __noinline__ float4 kpf_direct_slice(const float* __restrict__ q, const float4* __restrict__ s4,
                                                const IdxT* __restrict__ idx, int ld, const char* __restrict__ xslab,
                                                const float4 kpk, float ext2, float inv_ext, int n, int Ns, int H, int Cin,
                                                int lane) {
    RowGeom<NH> g;
    load_row_geom<IdxT, NH>(q, s4, idx, ld, n, Ns, H, Cin * 4, lane, g);
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
            const unsigned ob = (unsigned)__shfl_sync(0xffffffffu, g.sio[j], src);
            if (lane < 16) {
                const float4 xr = __ldg(reinterpret_cast<const float4*>(xslab + ob) + lane);
                acc.x = fmaf(wb, xr.x, acc.x); acc.y = fmaf(wb, xr.y, acc.y);
                acc.z = fmaf(wb, xr.z, acc.z); acc.w = fmaf(wb, xr.w, acc.w);
            }
        }
    }
    return acc;
}

template <typename IdxT, int CIN, int COUT, int NH>
__global__ void __launch_bounds__(KPF_THREADS, 1)
kpconv_fused_kernel(const __grid_constant__ CUtensorMap tmB, const float* __restrict__ q, const float4* __restrict__ s4,
                    const IdxT* __restrict__ idx, int ld, const float* __restrict__ x, const float* __restrict__ kp,
                    float extent, int Nq, int Ns, int H, int K, float* __restrict__ out, float* __restrict__ gstat) {
    using Cfg = KpfCfg<COUT>;
    constexpr int NSB = Cfg::NSB;
    constexpr int SLABS = CIN / 64;                                // 64-channel slabs per kernel point
    constexpr int KB_PER_K = CIN / GEMM_BK;                        // 32-channel k-blocks per kernel point
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* const gen = smem_raw + (base - raw);                  // generic pointer to the aligned base
    const uint32_t sA = base, sB = base + KPF_NSA * Cfg::A_BYTES;
    const uint32_t oL = KPF_NSA * Cfg::A_BYTES + NSB * Cfg::B_BYTES;         // lists
    const uint32_t oN = oL + Cfg::LISTS;                                     // inv_nn[128]
    const uint32_t bars = base + oN + 512;
    // fullA[NSA] (count NPW), emptyA[NSA], fullB[NSB], emptyB[NSB], tmem_full
    const uint32_t bar_fullA = bars, bar_emptyA = bars + 8 * KPF_NSA, bar_fullB = bars + 16 * KPF_NSA,
                   bar_emptyB = bar_fullB + 8 * NSB, bar_tmem = bar_emptyB + 8 * NSB;
    __shared__ uint32_t s_tmem_base;
    __shared__ float4 s_kp[KP_MAX_K];
    float* s_invnn = reinterpret_cast<float*>(gen + oN);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * GEMM_BM;
    const int num_kb = K * KB_PER_K;

    if (threadIdx.x < KP_MAX_K)
        s_kp[threadIdx.x] = threadIdx.x < K ? make_float4(kp[3 * threadIdx.x], kp[3 * threadIdx.x + 1], kp[3 * threadIdx.x + 2], 0.f)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == KPF_NPW && lane == 0) {
        for (int s = 0; s < KPF_NSA; ++s) { mbar_init(bar_fullA + 8 * s, KPF_NPW); mbar_init(bar_emptyA + 8 * s, 1); }
        for (int s = 0; s < NSB; ++s) { mbar_init(bar_fullB + 8 * s, 1); mbar_init(bar_emptyB + 8 * s, 1); }
        mbar_init(bar_tmem, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == KPF_NPW + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(COUT));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == KPF_NPW) {
        if (lane == 0) {                                             // ===== TMA: weight k-blocks =====
            int s = 0; uint32_t ph = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(bar_emptyB + 8 * s, ph ^ 1);
                mbar_arrive_expect_tx(bar_fullB + 8 * s, Cfg::B_BYTES);
                tma_load_2d(sB + s * Cfg::B_BYTES, &tmB, bar_fullB + 8 * s, kb * GEMM_BK, 0);
                if (++s == NSB) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == KPF_NPW + 1) {
        if (lane == 0) {                                             // ===== MMA issuer =====
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(GEMM_BM >> 4) << 24);
            int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(bar_fullB + 8 * sb, phb);
                mbar_wait(bar_fullA + 8 * sa, pha);
                tc_fence_after();
                const uint64_t da = make_smem_desc(sA + sa * Cfg::A_BYTES), db = make_smem_desc(sB + sb * Cfg::B_BYTES);
#pragma unroll
                for (int k4 = 0; k4 < GEMM_BK / 8; ++k4)
                    tc_mma_tf32(tmem_base, da + 2 * k4, db + 2 * k4, idesc, (kb | k4) != 0);
                tc_commit(bar_emptyA + 8 * sa);
                tc_commit(bar_emptyB + 8 * sb);
                if (++sa == KPF_NSA) { sa = 0; pha ^= 1; }
                if (++sb == NSB) { sb = 0; phb ^= 1; }
            }
            tc_commit(bar_tmem);
        }
    } else {                                                         // ===== producers (warps 0..NPW-1) =====
        const float ext2 = extent * extent, inv_ext = 1.0f / extent;
        uint8_t* wl = gen + oL + (size_t)warp * 8 * KPF_LROW;        // this warp's 8 list slots
        const int rowbase = warp * 8;                                // tile rows rowbase .. rowbase+7
        // ---- phase 1: 4 rows at a time, all index loads, then all packed-record gathers, then the lists (the 16 warps
        // of the CTA are all the latency hiding there is: independent loads have to be in flight together) ----
        unsigned ovf_mask = 0;                                       // bit r: row r keeps no list (direct path)
#pragma unroll 1
        for (int r4 = 0; r4 < 8; r4 += 4) {
            int si[4][NH];
            float4 p[4][NH];
            float3 qq[4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int n = m0 + rowbase + r4 + rr;
#pragma unroll
                for (int j = 0; j < NH; ++j) {
                    const int h = j * 32 + lane;
                    long long v = Ns;
                    if (n < Nq && h < H) v = (long long)idx[(size_t)n * ld + h];
                    si[rr][j] = (v >= 0 && v < Ns) ? (int)v : Ns;
                }
                const int nc = min(n, Nq - 1);
                qq[rr] = make_float3(q[3 * (size_t)nc], q[3 * (size_t)nc + 1], q[3 * (size_t)nc + 2]);
            }
#pragma unroll
            for (int rr = 0; rr < 4; ++rr)
#pragma unroll
                for (int j = 0; j < NH; ++j) {
                    p[rr][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (si[rr][j] < Ns) p[rr][j] = __ldg(s4 + si[rr][j]);
                }
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int r = r4 + rr;
                uint32_t* otab = reinterpret_cast<uint32_t*>(wl + (size_t)r * KPF_LROW);
                uint32_t* ent = otab + 64;
                int* off = reinterpret_cast<int*>(ent + KPF_ECAP);
                RowGeom<NH> g;
                int nn = 0;
#pragma unroll
                for (int j = 0; j < NH; ++j) {
                    const bool valid = si[rr][j] < Ns;
                    nn += (valid && p[rr][j].w > 0.f) ? 1 : 0;
                    g.sio[j] = valid ? si[rr][j] * CIN * 4 : 0;
                    g.rx[j] = valid ? p[rr][j].x - qq[rr].x : 3e18f;
                    g.ry[j] = valid ? p[rr][j].y - qq[rr].y : 0.f;
                    g.rz[j] = valid ? p[rr][j].z - qq[rr].z : 0.f;
                    g.any[j] = __ballot_sync(0xffffffffu, valid);
                }
                nn = __reduce_add_sync(0xffffffffu, nn);
                if (lane == 0) s_invnn[rowbase + r] = 1.0f / (float)max(nn, 1);
                build_row_list_packed<NH>(g, s_kp, K, ext2, inv_ext, otab, ent, off, lane);
                __syncwarp();
                if (off[KP_MAX_K] > KPF_ECAP) ovf_mask |= 1u << r;
            }
        }
        __syncwarp();
        // ---- phase 2 ----
        const int grp = lane >> 4, lg = lane & 15;                   // lane group (row of the pair), float4 index in the slab
        const int stage_of_lg = lg >> 3;                             // which of the pass's two A stages (32-channel k-block)
        const int chunk = lg & 7;                                    // 16-byte chunk inside the 128-byte row
        const char* xb = reinterpret_cast<const char*>(x);
        int sa = 0; uint32_t pha = 0;                                // first A stage of the current pass, its phase
#pragma unroll 1
        for (int k = 0; k < K; ++k) {
#pragma unroll 1
            for (int slab = 0; slab < SLABS; ++slab) {
                if (lane == 0) { mbar_wait(bar_emptyA + 8 * sa, pha ^ 1); mbar_wait(bar_emptyA + 8 * (sa + 1), pha ^ 1); }
                __syncwarp();
                const char* xslab = xb + (size_t)slab * 256;         // + entry byte offset + lg * 16
                uint8_t* astage = gen + (size_t)(sa + stage_of_lg) * Cfg::A_BYTES;
                // the lane group's four rows (one per row pair) advance together: four independent gathers in flight
                const uint8_t* slot0 = wl + (size_t)grp * KPF_LROW;   // row slots grp, grp+2, grp+4, grp+6
                int e[4], end[4];
                float4 acc[4];
#pragma unroll
                for (int rp = 0; rp < 4; ++rp) {
                    const int* off = reinterpret_cast<const int*>(slot0 + (size_t)rp * 2 * KPF_LROW + 256 + KPF_ECAP * 4);
                    const bool listed = ((ovf_mask >> (rp * 2 + grp)) & 1u) == 0;
                    e[rp] = listed ? off[k] : 0;
                    end[rp] = listed ? off[k + 1] : 0;
                    acc[rp] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const int maxlen = max(max(end[0] - e[0], end[1] - e[1]), max(end[2] - e[2], end[3] - e[3]));
                const char* xl = xslab + lg * 16;
#pragma unroll 1
                for (int it = 0; it < maxlen; ++it) {
                    uint32_t en[4];
                    float4 xv[4];
#pragma unroll
                    for (int rp = 0; rp < 4; ++rp) {
                        const uint32_t* otab = reinterpret_cast<const uint32_t*>(slot0 + (size_t)rp * 2 * KPF_LROW);
                        en[rp] = 0u;
                        xv[rp] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (e[rp] < end[rp]) {
                            en[rp] = otab[64 + e[rp]];
                            xv[rp] = __ldg(reinterpret_cast<const float4*>(xl + otab[en[rp] & 63u]));
                        }
                        ++e[rp];
                    }
#pragma unroll
                    for (int rp = 0; rp < 4; ++rp) {
                        const float w = __uint_as_float(en[rp] & ~63u);       // 0 for finished lists
                        acc[rp].x = fmaf(w, xv[rp].x, acc[rp].x); acc[rp].y = fmaf(w, xv[rp].y, acc[rp].y);
                        acc[rp].z = fmaf(w, xv[rp].z, acc[rp].z); acc[rp].w = fmaf(w, xv[rp].w, acc[rp].w);
                    }
                }
                __syncwarp();
                if (ovf_mask) {                                      // rows that keep no list: whole-warp direct evaluation (rare)
#pragma unroll
                    for (int rp = 0; rp < 4; ++rp) {
#pragma unroll 1
                        for (int g2 = 0; g2 < 2; ++g2) {
                            if ((ovf_mask >> (rp * 2 + g2)) & 1u) {
                                const float4 d = kpf_direct_slice<IdxT, NH>(q, s4, idx, ld, xslab, s_kp[k], ext2, inv_ext,
                                                                            m0 + rowbase + rp * 2 + g2, Ns, H, CIN, lane);
                                const float4 dd = make_float4(__shfl_sync(0xffffffffu, d.x, lg), __shfl_sync(0xffffffffu, d.y, lg),
                                                              __shfl_sync(0xffffffffu, d.z, lg), __shfl_sync(0xffffffffu, d.w, lg));
                                if (grp == g2) acc[rp] = dd;
                            }
                        }
                    }
                }
                // store the row slices: SWIZZLE_128B K-major tile, row pitch 128 B, chunk j of row t at j ^ (t & 7)
#pragma unroll
                for (int rp = 0; rp < 4; ++rp) {
                    const int t = rowbase + rp * 2 + grp;
                    const float4 o = make_float4(pre_round_tf32(acc[rp].x), pre_round_tf32(acc[rp].y), pre_round_tf32(acc[rp].z), pre_round_tf32(acc[rp].w));
                    *reinterpret_cast<float4*>(astage + (size_t)t * 128 + ((chunk ^ (t & 7)) << 4)) = o;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) { mbar_arrive(bar_fullA + 8 * sa); mbar_arrive(bar_fullA + 8 * (sa + 1)); }
                sa += 2;
                if (sa == KPF_NSA) { sa = 0; pha ^= 1; }
            }
        }
        // ---- epilogue: warps 0..3 drain TMEM (lane quarter = warp) ----
        if (warp < 4) {
            if (lane == 0) mbar_wait(bar_tmem, 0);
            __syncwarp();
            tc_fence_after();
            const int quarter = warp;
            const int row_l = quarter * 32 + lane;
            const float sc = s_invnn[row_l];
            float* tbuf = reinterpret_cast<float*>(gen) + (size_t)warp * 32 * 36;   // aliases the A ring (all MMAs retired)
            const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
            const bool whole = m0 + quarter * 32 + 32 <= Nq;
#pragma unroll 1
            for (int c = 0; c < COUT; c += 32) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, v);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(tbuf + lane * 36 + j) = make_float4(__uint_as_float(v[j]) * sc, __uint_as_float(v[j + 1]) * sc,
                                                                                  __uint_as_float(v[j + 2]) * sc, __uint_as_float(v[j + 3]) * sc);
                __syncwarp();
#pragma unroll
                for (int r4 = 0; r4 < 32; r4 += 4) {
                    const int rr = r4 + sub_r, grow = m0 + quarter * 32 + rr;
                    if (grow < Nq)
                        *reinterpret_cast<float4*>(out + (size_t)grow * COUT + c + sub_c) = *reinterpret_cast<const float4*>(tbuf + rr * 36 + sub_c);
                }
                if (gstat && whole) {                                // group statistics for the InstanceNorm that follows
                    float sum = 0.f;
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) sum += tbuf[rr * 36 + lane];
                    const float mean = sum * (1.0f / 32.0f);
                    float m2 = 0.f;
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr) { const float d = tbuf[rr * 36 + lane] - mean; m2 = fmaf(d, d, m2); }
                    float* gp = gstat + (size_t)((m0 >> 5) + quarter) * 2 * COUT + c + lane;
                    gp[0] = mean; gp[COUT] = m2;
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == KPF_NPW + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(COUT));
    }
}

template <typename IdxT, int CIN, int COUT, int NH>
static int launch_fused(const CUtensorMap& tmB, const float* d_q, const float4* s4, const void* d_idx, int ld,
                        const float* d_x, const float* d_kp, float extent, int Nq, int Ns, int H, int K, float* d_out,
                        float* d_gstat, cudaStream_t st) {
    static bool attr_set[64] = {};                              // function attributes are per device
    const int dev_i = current_device() & 63;
    if (!attr_set[dev_i]) {
        APRB_CUDA_OK(cudaFuncSetAttribute(kpconv_fused_kernel<IdxT, CIN, COUT, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, KpfCfg<COUT>::SMEM));
        attr_set[dev_i] = true;
    }
    {
        ProfScope ps("kpconv_fused_kernel", st, 1);
        kpconv_fused_kernel<IdxT, CIN, COUT, NH><<<cdiv(Nq, GEMM_BM), KPF_THREADS, KpfCfg<COUT>::SMEM, st>>>(
            tmB, d_q, s4, (const IdxT*)d_idx, ld, d_x, d_kp, extent, Nq, Ns, H, K, d_out, d_gstat);
    }
    APRB_LAUNCH_OK();
    return APRB_OK;
}

// aprb_set_option("kpconv_fused"): 1 routes Cin = Cout in {64, 128}, H <= 64 through the one-kernel KPConv. Parity-green
// but measured 2.2x slower than the two-kernel path on B200 (profiles/r01_kpconv_fused.txt), so off by default.
int g_kpconv_fused = 0;

bool kpconv_fused_supported(int H, int K, int Cin, int Cout, long long Ns) {
    return g_kpconv_fused && H <= 64 && K >= 1 && K <= KP_MAX_K && ((Cin == 64 && Cout == 64) || (Cin == 128 && Cout == 128)) &&
           Ns * Cin < (1LL << 30);
}

// s4: packed support records (x, y, z, flag), already built by the caller (rowsum_pos_kernel, kpconv.cu)
int kpconv_fused_run(const float* d_q, const float4* s4, const void* d_idx, int idx_is_i64, int ld, const float* d_x,
                     const float* d_kp, const float* d_wprep, float extent, int Nq, int Ns, int H, int K, int Cin, int Cout,
                     float* d_out, float* d_gstat, cudaStream_t st) {
    CUtensorMap tmB;
    int rc = make_tmap(&tmB, d_wprep, Cout, K * Cin, Cout);
    if (rc) return rc;
#define KPF_GO(IDX, CI, CO)                                                                                          \
    (H <= 32 ? launch_fused<IDX, CI, CO, 1>(tmB, d_q, s4, d_idx, ld, d_x, d_kp, extent, Nq, Ns, H, K, d_out, d_gstat, st) \
             : launch_fused<IDX, CI, CO, 2>(tmB, d_q, s4, d_idx, ld, d_x, d_kp, extent, Nq, Ns, H, K, d_out, d_gstat, st))
    if (Cin == 64) return idx_is_i64 ? KPF_GO(long long, 64, 64) : KPF_GO(int, 64, 64);
    return idx_is_i64 ? KPF_GO(long long, 128, 128) : KPF_GO(int, 128, 128);
#undef KPF_GO
}

}  // namespace aprb
