// kpconv_tc.cu — K5 stage A+B (kernel-point weighting) on the 5th-generation tensor cores.
//
// Replaces the weighting half of KPConv.forward (/root/reference/Predator_APR/models/blocks.py:269-354):
//   w[n,k,h]  = max(0, 1 - |s[idx[n,h]] - q[n] - kp[k]| / extent)          (:269-289, :328-329)
//   wf[n,k,:] = sum_h w[n,k,h] * x[idx[n,h],:]                              (:348-354, torch.matmul(all_weights, neighb_x))
// The reference evaluates this as a dense per-query product [K x H] @ [H x Cin]. tcgen05.mma needs M >= 64 rows that share
// one B operand, which per-query neighbourhoods do not offer when the query is put on M — so the product is issued
// TRANSPOSED: M = channels (64 or 128 lanes of TMEM), N = 16 kernel points (15 + one zero column), K = the query's
// neighbours (up to 64, in steps of 16):
//   D1[c, k] = sum_h  X_q[c, h] * w_q[h, k]
// A = the gathered feature rows of the query, which are exactly an MN-major (channel-contiguous) SWIZZLE_128B operand when
//     each neighbour's row is dropped into shared memory as one 128-byte line per 64 channels (cp.async, 16-byte chunk j of
//     row r at chunk j ^ (r & 7)): no transposition, no per-element work;
// B = the influence weights of the query, K-major SWIZZLE_128B [16 kernel points x 64 neighbours] fp16, computed by the
//     producer warp (lanes = neighbour pairs) and stored as half2.
// One MMA per (query, 16 neighbours, 128 channels); the accumulator D1 (16 TMEM columns per query) is drained by four
// epilogue warps (lane = channel) which round it to fp16 and store it as wf[n, c, 0..15] (channel-major, kernel point
// minor: the 16 values a lane holds are contiguous, so a warp writes 1 KB with two 128-bit stores per lane) — the A
// operand of the [Nq, 16*Cin] x [16*Cin, Cout] contraction (gemm_tcgen05.cu) against weights prepared in the same
// order (aprb_kpconv_prepare_weights_f16_ck; the 16th column / row is zero). Per query this costs ~450 warp instructions (the 15 x 64 influence evaluations) instead of ~1660, and the
// 17 GFLOP/pair of the weighting stage run on the tensor pipe.
//
// The unit of the CTA's pipeline is a TILE of TQ consecutive queries (8 / 6 / 3 for Cin = 64 / 128 / 256), so that every
// hand-off (mbarrier test ~100-300 cycles, tcgen05.commit, TMEM drain) is paid once per tile, not once per query — with a
// per-query pipeline the four epilogue warps, which had to touch every query in turn, ran at ~900 cycles per query and
// everything else queued behind them (r02 event trace, profiles/r02_kpconv_tc.txt).
//   producers   TQ teams of 3 warps, one query of the tile each: warp 0 gathers the feature rows (index row re-read from a
//               256-byte shared copy: one LDS + one address + one cp.async moves 512 bytes; pads are zero-filled through
//               cp.async's src-size operand), warps 1 and 2 evaluate 8 kernel points each; all prefetch the next tile's
//               indices / support records. Two tile slots in shared memory (double buffer).
//   MMA         4 issuer warps split the queries of a tile; the loop is warp-uniform (only the tcgen05 instructions are
//               predicated to lane 0) so that descriptors live in uniform registers.
//   epilogue    1 or 2 groups of 4 warps (TMEM lane quarter = warp & 3) drain a tile's accumulators from one of two TMEM
//               buffers. With Cin = 64 the accumulator of a query occupies 16 lanes of every 32-lane TMEM quarter (M = 64),
//               so two consecutive queries interleave in one 16-column slot (lane offset 16) and are drained together.
#include "kpconv_common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace aprb {

constexpr int KTC_NI = 4;                                      // MMA issuer warps
constexpr int KTC_TW = 3;                                      // warps per producer team: gather | kernel points 0-7 | 8-15
constexpr int KTC_TCOLS = 256;                                 // TMEM columns allocated: two buffers of 128

template <int CIN>
struct KtcCfg {
    static constexpr int MB = (CIN + 127) / 128;               // MMAs per K step (128 channels each)
    static constexpr int MM = CIN >= 128 ? 128 : 64;           // M of one MMA
    static constexpr int A_BYTES = CIN * 128;                  // CIN/64 atoms of [64 neighbour rows][128 B]
    static constexpr int B_BYTES = 2048;                       // [16 kernel points][64 neighbours] fp16
    static constexpr int SLOT = A_BYTES + B_BYTES;             // one query
    static constexpr int TQ = CIN <= 64 ? 8 : (CIN <= 128 ? 6 : 3);   // queries per tile = producer teams
    static constexpr int QPT = CIN == 64 ? 2 : 1;              // queries per 16*MB-column TMEM slot
    static constexpr int CS = TQ / QPT;                        // TMEM column slots per tile
    static constexpr int EG = CIN == 64 ? 1 : 2;               // epilogue groups of 4 warps
    static constexpr int EPI_WARPS = 4 * EG;
    static constexpr int PROD0 = EPI_WARPS + KTC_NI;
    static constexpr int THREADS = (PROD0 + KTC_TW * TQ) * 32;
    static constexpr int TILE_BYTES = TQ * SLOT;
    static constexpr int TAIL = 1024 + 2 * TQ * 256;           // barriers + K steps, then the index rows of both tile slots
    static constexpr int SMEM = 2 * TILE_BYTES + 1024 /*align*/ + TAIL;
    static_assert(CS * 16 * MB <= 128, "a tile's accumulators must fit one TMEM buffer");
    static_assert(TQ % QPT == 0, "pairs must not straddle tiles");
};

// 16-byte global -> shared copy; src_bytes = 0 zero-fills the destination without reading (shadow neighbours)
template <bool L1>
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
    if (L1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MN-major (M contiguous), SWIZZLE_128B operand descriptor (cute/arch/mma_sm100_desc.hpp, make_umma_desc<Major::MN>):
// canonical layout in 16-byte units ((8, n), (8, k)) : ((1, LBO), (8, SBO)) — 64 fp16 of M contiguous per K row, the next
// 64 of M at +LBO bytes, the next group of 8 K rows at +SBO bytes.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

template <int CIN, bool L1>
__global__ void __launch_bounds__(KtcCfg<CIN>::THREADS, 1)
kpconv_tc_kernel(const float* __restrict__ q, const float4* __restrict__ s4, const int* __restrict__ idx, int ld,
                 const __half* __restrict__ x, const float* __restrict__ kp, float extent, int Nq, int Ns, int H, int K,
                 int per_cta, __half* __restrict__ wf, float* __restrict__ inv_nn, int dbg) {
    using Cfg = KtcCfg<CIN>;
    constexpr int MB = Cfg::MB, QPT = Cfg::QPT, TQ = Cfg::TQ, CS = Cfg::CS, EG = Cfg::EG;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* const gen = smem_raw + (base - raw);
    const uint32_t bars = base + 2 * Cfg::TILE_BYTES;
    // per tile slot (shared memory): full (producers -> MMA), empty (MMA -> producers); per TMEM buffer: dfull, dempty
    const uint32_t bar_full = bars, bar_empty = bars + 16, bar_dfull = bars + 32, bar_dempty = bars + 48;
    int* const s_kr = reinterpret_cast<int*>(gen + 2 * Cfg::TILE_BYTES + 64);            // [2][TQ] K steps of every query (0 = none)
    uint32_t* const s_idx = reinterpret_cast<uint32_t*>(gen + 2 * Cfg::TILE_BYTES + 1024);   // [2][TQ][64] index rows
    __shared__ uint32_t s_tmem_base;
    __shared__ float4 s_kp[KP_MAX_K];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * per_cta;
    const int nq = min(per_cta, Nq - q0);                        // queries of this CTA (> 0 by the launch)
    const int ntiles = (nq + TQ - 1) / TQ;

    if (threadIdx.x < KP_MAX_K)                                  // unused kernel points sit far away on y: weight 0
        s_kp[threadIdx.x] = threadIdx.x < K ? make_float4(kp[3 * threadIdx.x], kp[3 * threadIdx.x + 1], kp[3 * threadIdx.x + 2], 0.f)
                                            : make_float4(0.f, 3e18f, 0.f, 0.f);
    if (warp == Cfg::EPI_WARPS && lane == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_full + 8 * b, KTC_TW * TQ); mbar_init(bar_empty + 8 * b, KTC_NI);
            mbar_init(bar_dfull + 8 * b, KTC_NI); mbar_init(bar_dempty + 8 * b, Cfg::EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == Cfg::EPI_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(KTC_TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp >= Cfg::PROD0) {
        const int team = (warp - Cfg::PROD0) / KTC_TW, tw = (warp - Cfg::PROD0) % KTC_TW;
        if (tw == 0) {
            // ================= gather warp: A = the query's feature rows =================
            constexpr int CPR = CIN / 8;                           // 16-byte chunks per feature row
            constexpr int RPI = CPR >= 32 ? 1 : 32 / CPR;          // rows per warp-wide copy instruction
            constexpr int IPR = CPR >= 32 ? CPR / 32 : 1;          // copy instructions per row
            constexpr int U = 8 / RPI;                             // copy steps per group of 8 rows
            const int rl = CPR >= 32 ? 0 : lane / CPR;             // this lane's row within a copy step
            const int j0 = CPR >= 32 ? lane : lane % CPR;          // this lane's first chunk within the row
            uint32_t pre[U][IPR];                                  // swizzled destination of (step u, chunk) within a row group
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int i = 0; i < IPR; ++i) {
                    const int r = rl + u * RPI, j = j0 + 32 * i;
                    pre[u][i] = (uint32_t)(j >> 3) * 8192u + (uint32_t)r * 128u + (uint32_t)(((j & 7) ^ (r & 7)) << 4);
                }
            const char* xlane = reinterpret_cast<const char*>(x) + j0 * 16;
            int v0n = Ns, v1n = Ns;                                // prefetched index row of this team's query in the next tile
            if (team < nq) {
                const size_t rowi = (size_t)(q0 + team) * ld;
                if (lane < H) v0n = idx[rowi + lane];
                if (lane + 32 < H) v1n = idx[rowi + 32 + lane];
            }
            for (int it = 0; it < ntiles; ++it) {
                const int qs = it * TQ + team, b = it & 1;
                const bool have = qs < nq;
                const uint32_t a_base = base + b * Cfg::TILE_BYTES + team * Cfg::SLOT;
                const int v0 = (have && v0n >= 0 && v0n < Ns) ? v0n : Ns, v1 = (have && v1n >= 0 && v1n < Ns) ? v1n : Ns;
                if (qs + TQ < nq) {                                  // next tile's index row: in flight during this gather
                    const size_t rowi = (size_t)(q0 + qs + TQ) * ld;
                    v0n = lane < H ? idx[rowi + lane] : Ns;
                    v1n = lane + 32 < H ? idx[rowi + 32 + lane] : Ns;
                }
                const unsigned m0 = __ballot_sync(0xffffffffu, v0 < Ns), m1 = __ballot_sync(0xffffffffu, v1 < Ns);
                const int last = m1 ? 64 - __clz(m1) : (m0 ? 32 - __clz(m0) : 0);          // 1 + last valid neighbour
                const int kr = have ? max((last + 15) >> 4, 1) : 0;   // K steps (16 neighbours each) the MMA has to cover
                if (lane == 0) mbar_wait(bar_empty + 8 * b, ((it >> 1) & 1) ^ 1);   // the MMAs that read this tile slot have retired
                __syncwarp();
                uint32_t* si = s_idx + (b * TQ + team) * 64;
                si[lane] = (uint32_t)v0; si[32 + lane] = (uint32_t)v1;
                __syncwarp();
                if (!(dbg & 1)) {
#pragma unroll 1
                    for (int g = 0; g < kr * 2; ++g) {               // groups of 8 neighbour rows
                        int sir[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) sir[u] = (int)si[g * 8 + u * RPI + rl];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const bool ok = sir[u] < Ns;
                            const char* src = xlane + (size_t)(ok ? sir[u] : 0) * (CIN * 2);
#pragma unroll
                            for (int i = 0; i < IPR; ++i)
                                cp_async16_zfill<L1>(a_base + (uint32_t)g * 1024u + pre[u][i], src + i * 512, ok ? 16u : 0u);
                        }
                    }
                }
                cp_async_wait_all();
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) { s_kr[b * TQ + team] = kr; mbar_arrive(bar_full + 8 * b); }
            }
        } else {
            // ================= weight warps: B = influence of 8 kernel points on the 64 neighbours =================
            const int k0 = (tw - 1) * 8;
            const float inv_ext = 1.0f / extent;
            // two-deep prefetch: indices of the query after next, support records of the next query
            int a0 = Ns, a1 = Ns, b0 = Ns, b1 = Ns;                 // a: current query's neighbours 2l, 2l+1; b: next query's
            float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
            float qx = 0.f, qy = 0.f, qz = 0.f;
            auto load_idx = [&](int qsi, int& i0, int& i1) {
                i0 = Ns; i1 = Ns;
                if (qsi < nq) {
                    const size_t rowi = (size_t)(q0 + qsi) * ld + 2 * lane;
                    if (2 * lane < H) { const int v = idx[rowi]; i0 = (v >= 0 && v < Ns) ? v : Ns; }
                    if (2 * lane + 1 < H) { const int v = idx[rowi + 1]; i1 = (v >= 0 && v < Ns) ? v : Ns; }
                }
            };
            load_idx(team, a0, a1);
            load_idx(team + TQ, b0, b1);
            if (team < nq) {
                if (a0 < Ns) p0 = __ldg(s4 + a0);
                if (a1 < Ns) p1 = __ldg(s4 + a1);
                qx = q[3 * (size_t)(q0 + team)]; qy = q[3 * (size_t)(q0 + team) + 1]; qz = q[3 * (size_t)(q0 + team) + 2];
            }
            for (int it = 0; it < ntiles; ++it) {
                const int qs = it * TQ + team, b = it & 1;
                const int n = q0 + qs;
                // current query: relative positions (shadow neighbours sit at x = 3e18: outside every extent)
                const float r0x = a0 < Ns ? p0.x - qx : 3e18f, r0y = p0.y - qy, r0z = p0.z - qz;
                const float r1x = a1 < Ns ? p1.x - qx : 3e18f, r1y = p1.y - qy, r1z = p1.z - qz;
                int nn = (a0 < Ns && p0.w > 0.f ? 1 : 0) + (a1 < Ns && p1.w > 0.f ? 1 : 0);
                // next tile's query: support records and query point now, indices of the one after it
                a0 = b0; a1 = b1;
                p0 = make_float4(0.f, 0.f, 0.f, 0.f); p1 = p0;
                if (qs + TQ < nq) {
                    if (a0 < Ns) p0 = __ldg(s4 + a0);
                    if (a1 < Ns) p1 = __ldg(s4 + a1);
                    const size_t nn3 = 3 * (size_t)(n + TQ);
                    qx = q[nn3]; qy = q[nn3 + 1]; qz = q[nn3 + 2];
                }
                load_idx(qs + 2 * TQ, b0, b1);
                // all 8 kernel points first (independent chains), then the 8 stores: row k of the K-major SWIZZLE_128B tile
                // is 128 bytes = the 64 neighbours; this lane's pair is 4-byte word (lane & 3) of chunk (lane >> 2) ^ (k & 7)
                uint32_t hw[8];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const float4 kpk = s_kp[k0 + kk];
                    float dx = r0x - kpk.x, dy = r0y - kpk.y, dz = r0z - kpk.z;
                    const float w0 = fmaxf(1.0f - sqrt_approx(dx * dx + dy * dy + dz * dz) * inv_ext, 0.f);
                    dx = r1x - kpk.x; dy = r1y - kpk.y; dz = r1z - kpk.z;
                    const float w1 = fmaxf(1.0f - sqrt_approx(dx * dx + dy * dy + dz * dz) * inv_ext, 0.f);
                    const __half2 t = __floats2half2_rn(w0, w1);
                    hw[kk] = *reinterpret_cast<const uint32_t*>(&t);
                }
                if (lane == 0) mbar_wait(bar_empty + 8 * b, ((it >> 1) & 1) ^ 1);
                __syncwarp();
                if (qs < nq) {
                    uint32_t* brow = reinterpret_cast<uint32_t*>(gen + (size_t)b * Cfg::TILE_BYTES + (size_t)team * Cfg::SLOT + Cfg::A_BYTES +
                                                                 (size_t)k0 * 128) + (lane & 3);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) brow[kk * 32 + (((lane >> 2) ^ kk) << 2)] = hw[kk];   // (k0 + kk) & 7 == kk
                }
                if (tw == 1) nn = __reduce_add_sync(0xffffffffu, nn);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (tw == 1 && qs < nq) inv_nn[n] = 1.0f / (float)max(nn, 1);
                    mbar_arrive(bar_full + 8 * b);
                }
            }
        }
    } else if (warp >= Cfg::EPI_WARPS) {
        // ================= MMA issuers: issuer mi runs the queries mi, mi + NI, ... of every tile =================
        // The loop is warp-uniform; only the tcgen05 instructions are predicated to lane 0.
        const int mi = warp - Cfg::EPI_WARPS;
        // D fp32, A/B fp16, A MN-major (bit 15), B K-major, N = 16, M = 64 / 128
        const uint32_t idesc = (1u << 4) | (1u << 15) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(Cfg::MM >> 4) << 24);
        for (int it = 0; it < ntiles; ++it) {
            const int b = it & 1;
            if (lane == 0) { mbar_wait(bar_full + 8 * b, (it >> 1) & 1); mbar_wait(bar_dempty + 8 * b, ((it >> 1) & 1) ^ 1); }
            __syncwarp();
            tc_fence_after();
#pragma unroll 1
            for (int j = mi; j < TQ; j += KTC_NI) {
                const int kr = (dbg & 2) ? min(1, s_kr[b * TQ + j]) : s_kr[b * TQ + j];
                const uint32_t a_base = base + b * Cfg::TILE_BYTES + j * Cfg::SLOT, b_base = a_base + Cfg::A_BYTES;
                const uint64_t db = make_smem_desc(b_base);
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                    const uint32_t tmem_d = tmem_base + (uint32_t)(b * 128 + (j / QPT) * 16 * MB + mb * 16) +
                                            (QPT == 2 ? ((uint32_t)((j & 1) * 16) << 16) : 0u);
                    for (int ks = 0; ks < kr; ++ks) {
                        const uint64_t da = make_smem_desc_mn(a_base + (uint32_t)mb * 16384u + (uint32_t)ks * 2048u, 8192u, 1024u);
                        if (lane == 0) tc_mma_f16(tmem_d, da, db + 2 * ks, idesc, ks != 0);
                    }
                }
            }
            if (lane == 0) { tc_commit(bar_empty + 8 * b); tc_commit(bar_dfull + 8 * b); }
            __syncwarp();
        }
    } else {
        // ================= epilogue: TMEM -> fp16 -> wf[n, c, k] =================
        const int e = warp & 3, g = warp >> 2;                        // TMEM lane quarter, epilogue group
        constexpr int SPG = (CS + EG - 1) / EG;                       // column slots per group and tile
        for (int it = 0; it < ntiles; ++it) {
            const int b = it & 1;
            if (lane == 0) mbar_wait(bar_dfull + 8 * b, (it >> 1) & 1);
            __syncwarp();
            tc_fence_after();
            constexpr int LDB = MB == 1 ? 2 : 1;                     // column slots drained per batch (32 accumulator registers)
#pragma unroll
            for (int i0 = 0; i0 < SPG; i0 += LDB) {
                uint32_t v[LDB][MB][16];
#pragma unroll
                for (int ii = 0; ii < LDB; ++ii) {
                    const int cs = g + (i0 + ii) * EG;
                    if (i0 + ii < SPG && cs < CS) {
#pragma unroll
                        for (int mb = 0; mb < MB; ++mb)
                            tc_ld16_nowait(tmem_base + ((uint32_t)(32 * e) << 16) + (uint32_t)(b * 128 + cs * 16 * MB + mb * 16), v[ii][mb]);
                    }
                }
                tc_ld_wait();
                if (i0 + LDB >= SPG) {                               // last batch: the TMEM buffer may be overwritten
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_dempty + 8 * b);
                }
#pragma unroll
                for (int ii = 0; ii < LDB; ++ii) {
                    const int cs = g + (i0 + ii) * EG;
                    if (i0 + ii >= SPG || cs >= CS) continue;
                    int n, c;
                    if (QPT == 2) { n = q0 + it * TQ + 2 * cs + (lane >> 4); c = 16 * e + (lane & 15); }
                    else { n = q0 + it * TQ + cs; c = 32 * e + lane; }
                    if (n < q0 + nq && !(dbg & 4)) {
                        // wf[n, c, 0..15]: channel-major, kernel point minor (columns K.. are zero: their weights are) — 32
                        // contiguous bytes per lane, 1 KB per warp
                        uint4* row = reinterpret_cast<uint4*>(wf + ((size_t)n * CIN + c) * KP_MAX_K);
#pragma unroll
                        for (int mb = 0; mb < MB; ++mb) {
                            uint32_t h[8];
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                h[k] = pack_half2_sat(__uint_as_float(v[ii][mb][2 * k]), __uint_as_float(v[ii][mb][2 * k + 1]));
                            }
                            row[mb * 128 * 2] = make_uint4(h[0], h[1], h[2], h[3]);
                            row[mb * 128 * 2 + 1] = make_uint4(h[4], h[5], h[6], h[7]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == Cfg::EPI_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(KTC_TCOLS));
    }
}

int g_ktc_dbg = 0;     // aprb_set_option("ktc_dbg"): diagnostic bits (1 no gather, 2 one K step, 4 no stores, 32 cp.async.ca)
// aprb_set_option("kpconv_tc"): weighting stage on tcgen05 (fp16 features, Cin in {64, 128, 256}, H <= 64). Parity-green
// (2.9e-4 vs the fp32 definition) but measured 1.15-1.7x SLOWER than the CUDA-core list kernel on B200 (r02:
// profiles/r02_kpconv_tc.txt: the 15 x 64 influence evaluations stay on CUDA cores and dominate), so off by default.
int g_kpconv_tc = 0;

bool kpconv_tc_supported(int H, int K, int Cin, long long Ns) {
    return g_kpconv_tc && H >= 1 && H <= 64 && K >= 1 && K <= KP_MAX_K && (Cin == 64 || Cin == 128 || Cin == 256) &&
           Ns * Cin < (1LL << 30);
}

template <int CIN, bool L1>
static int launch_ktc(const float* d_q, const float4* s4, const int* d_idx, int ld, const void* d_x16, const float* d_kp,
                      float extent, int Nq, int Ns, int H, int K, void* d_wf16, float* d_inv_nn, cudaStream_t st) {
    using Cfg = KtcCfg<CIN>;
    static bool attr_set[64] = {};
    const int dev_i = current_device() & 63;
    if (!attr_set[dev_i]) {
        APRB_CUDA_OK(cudaFuncSetAttribute(kpconv_tc_kernel<CIN, L1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_set[dev_i] = true;
    }
    int per = cdiv(Nq, sm_count());
    per = cdiv(per, Cfg::TQ) * Cfg::TQ;                              // whole tiles (and whole pairs) per CTA
    const int grid = cdiv(Nq, per);
    {
        ProfScope ps("kpconv_tc_kernel", st, 1);
        kpconv_tc_kernel<CIN, L1><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(d_q, s4, d_idx, ld, (const __half*)d_x16, d_kp, extent, Nq,
                                                                       Ns, H, K, per, (__half*)d_wf16, d_inv_nn, g_ktc_dbg);
    }
    APRB_LAUNCH_OK();
    return APRB_OK;
}

// Weighted tile wf [Nq, Cin*16] fp16 (channel-major, kernel point minor) and inv_nn [Nq] from fp16 features; s4 = packed
// support records (x, y, z, flag).
int kpconv_tc_run(const float* d_q, const float4* s4, const int* d_idx, int ld, const void* d_x16, const float* d_kp,
                  float extent, int Nq, int Ns, int H, int K, int Cin, void* d_wf16, float* d_inv_nn, cudaStream_t st) {
#define KTC_GO(C)                                                                                                            \
    ((g_ktc_dbg & 32) ? launch_ktc<C, true>(d_q, s4, d_idx, ld, d_x16, d_kp, extent, Nq, Ns, H, K, d_wf16, d_inv_nn, st)     \
                      : launch_ktc<C, false>(d_q, s4, d_idx, ld, d_x16, d_kp, extent, Nq, Ns, H, K, d_wf16, d_inv_nn, st))
    if (Cin == 64) return KTC_GO(64);
    if (Cin == 128) return KTC_GO(128);
    return KTC_GO(256);
#undef KTC_GO
}

}  // namespace aprb
