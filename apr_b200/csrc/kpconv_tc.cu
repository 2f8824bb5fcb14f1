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
// CTA = 4 epilogue warps + 1 MMA warp + up to 8 producer TEAMS of 3 warps, persistent over a contiguous range of queries.
// A team fills one shared-memory slot per query: warp 0 of the team gathers the feature rows (its lanes hold the index
// row; the rows' indices are re-read from a 256-byte shared copy so that one LDS + one address + one cp.async moves 512
// bytes; pads are zero-filled by cp.async's src-size operand), warps 1 and 2 evaluate the influence of 8 kernel points
// each. All three prefetch the next query's indices / support records while they work on the current one. A ring of NS
// slots sits between the teams and the MMA thread, a ring of NT TMEM slots between the MMA thread and the epilogue. With
// Cin = 64 the accumulator of a query occupies 16 lanes of every 32-lane TMEM quarter (M = 64), so two consecutive
// queries interleave in one slot (lane offset 16) and the epilogue drains them together.
#include "kpconv_common.cuh"
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace aprb {

constexpr int KTC_EPI_WARPS = 4;                               // warps 0-3: TMEM lane quarter = warp index
constexpr int KTC_MMA_WARP = 4;                                // first MMA warp
constexpr int KTC_MMA_WARPS = 4;                               // MMA issuers (lane 0 of each): TMEM slot ts goes to issuer ts % 4 —
                                                               // per query the issuing thread spends ~800 cycles in barrier
                                                               // tests, descriptor arithmetic and the issue latency of four
                                                               // small MMAs; one issuer alone was the bottleneck of the kernel
constexpr int KTC_PROD0 = KTC_MMA_WARP + KTC_MMA_WARPS;
constexpr int KTC_TW = 3;                                      // warps per producer team: gather | kernel points 0-7 | 8-15
constexpr int KTC_MAX_TEAMS = 8;
constexpr int KTC_TCOLS = 256;                                 // TMEM columns allocated

template <int CIN>
struct KtcCfg {
    static constexpr int MB = (CIN + 127) / 128;               // MMAs per K step (128 channels each)
    static constexpr int MM = CIN >= 128 ? 128 : 64;           // M of one MMA
    static constexpr int A_BYTES = CIN * 128;                  // CIN/64 atoms of [64 neighbour rows][128 B]
    static constexpr int B_BYTES = 2048;                       // [16 kernel points][64 neighbours] fp16
    static constexpr int SLOT = A_BYTES + B_BYTES;
    static constexpr int NS = CIN <= 64 ? 16 : (CIN <= 128 ? 12 : 6);
    static constexpr int NT = CIN == 64 ? KTC_TCOLS / 16 : KTC_TCOLS / (16 * MB);
    static constexpr int QPT = CIN == 64 ? 2 : 1;              // queries per TMEM slot
    // Teams: every team OWNS two slots and alternates between them (slot = 2 * team + parity of its query count). A team
    // waits for "the MMAs that read this slot" by mbarrier phase parity, which is only unambiguous while the waiter is at
    // most one phase ahead of the barrier. With a private slot pair that holds by construction — the previous user of the
    // slot is the team's own query before last, whose release it has itself waited for — whichever of the four MMA issuers
    // ran it and however far the issuers drift apart.
    static constexpr int NTEAM = NS / 2;
    static constexpr int THREADS = (KTC_PROD0 + KTC_TW * NTEAM) * 32;
    static constexpr int TAIL = 1024 + NS * 256;               // barriers + K steps, then the per-slot index rows
    static constexpr int SMEM = NS * SLOT + 1024 /*align*/ + TAIL;
};

// 16-byte global -> shared copy; src_bytes = 0 zero-fills the destination without reading (shadow neighbours)
template <bool L1>
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
    if (L1) asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void sts32(uint32_t dst, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t src) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(src) : "memory");
    return v;
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Wait with back-off: a waiter that polls (try_wait + loop control, ~8 instructions per ~100 cycles) costs issue slots that
// the working warps of this issue-bound kernel need; sleeping between tests trades a little wake-up latency for them.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, unsigned ns) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
        __nanosleep(ns);
        if (spins > (1u << 24)) __trap();
    }
}

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
                 int per_cta, __half* __restrict__ wf, float* __restrict__ inv_nn, int dbg, long long* __restrict__ prof) {
#define KTC_T(var) const long long var = (dbg & 16) ? clock64() : 0
#define KTC_EV(qsi, e) do { if ((dbg & 16) && blockIdx.x == 0 && (qsi) < 96 && prof) prof[128 + (qsi) * 8 + (e)] = clock64(); } while (0)
#define KTC_WAIT(bar, par) do { if (dbg & 64) mbar_wait_sleep(bar, par, dbg >> 8); else mbar_wait(bar, par); } while (0)
    using Cfg = KtcCfg<CIN>;
    constexpr int NS = Cfg::NS, NT = Cfg::NT, MB = Cfg::MB, QPT = Cfg::QPT, NTEAM = Cfg::NTEAM;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* const gen = smem_raw + (base - raw);
    const uint32_t bars = base + NS * Cfg::SLOT;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * NS, bar_dfull = bars + 16 * NS, bar_dempty = bar_dfull + 8 * NT;
    int* const s_kr = reinterpret_cast<int*>(gen + NS * Cfg::SLOT + 16 * NS + 16 * NT);   // K steps of the query in each slot
    const uint32_t sidx = bars + 1024;                            // [NS][64] int: the index row of the query in each slot
    __shared__ uint32_t s_tmem_base;
    __shared__ float4 s_kp[KP_MAX_K];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * per_cta;
    const int nq = min(per_cta, Nq - q0);                        // queries of this CTA (> 0 by the launch)

    if (threadIdx.x < KP_MAX_K)                                  // unused kernel points sit far away on y: weight 0
        s_kp[threadIdx.x] = threadIdx.x < K ? make_float4(kp[3 * threadIdx.x], kp[3 * threadIdx.x + 1], kp[3 * threadIdx.x + 2], 0.f)
                                            : make_float4(0.f, 3e18f, 0.f, 0.f);
    if (warp == KTC_MMA_WARP && lane == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(bar_full + 8 * s, KTC_TW); mbar_init(bar_empty + 8 * s, 1); }
        for (int t = 0; t < NT; ++t) { mbar_init(bar_dfull + 8 * t, QPT); mbar_init(bar_dempty + 8 * t, KTC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == KTC_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(KTC_TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp >= KTC_PROD0) {
        const int team = (warp - KTC_PROD0) / KTC_TW, tw = (warp - KTC_PROD0) % KTC_TW;
        long long pt[4] = {0, 0, 0, 0};
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
            int v0n = Ns, v1n = Ns;                                // prefetched index row of the next query
            if (team < nq) {
                const size_t rowi = (size_t)(q0 + team) * ld;
                if (lane < H) v0n = idx[rowi + lane];
                if (lane + 32 < H) v1n = idx[rowi + 32 + lane];
            }
            for (int qs = team; qs < nq; qs += NTEAM) {
                KTC_T(t0);
                const int s = qs % NS;
                const uint32_t a_base = base + s * Cfg::SLOT;
                const int v0 = (v0n >= 0 && v0n < Ns) ? v0n : Ns, v1 = (v1n >= 0 && v1n < Ns) ? v1n : Ns;
                if (qs + NTEAM < nq) {                               // next query's index row: in flight during this gather
                    const size_t rowi = (size_t)(q0 + qs + NTEAM) * ld;
                    v0n = lane < H ? idx[rowi + lane] : Ns;
                    v1n = lane + 32 < H ? idx[rowi + 32 + lane] : Ns;
                }
                const unsigned m0 = __ballot_sync(0xffffffffu, v0 < Ns), m1 = __ballot_sync(0xffffffffu, v1 < Ns);
                const int last = m1 ? 64 - __clz(m1) : (m0 ? 32 - __clz(m0) : 0);          // 1 + last valid neighbour
                const int kr = max((last + 15) >> 4, 1);             // K steps (16 neighbours each) the MMA has to cover
                KTC_T(t1);
                if (lane == 0) KTC_WAIT(bar_empty + 8 * s, ((qs / NS) & 1) ^ 1);   // the MMAs that read this slot have retired
                __syncwarp();
                KTC_T(t2);
                const uint32_t si_base = sidx + (uint32_t)s * 256u;
                sts32(si_base + lane * 4, (uint32_t)v0);
                sts32(si_base + 128 + lane * 4, (uint32_t)v1);
                __syncwarp();
                const uint32_t si_lane = si_base + (uint32_t)rl * 4u;
                if (!(dbg & 1)) {
#pragma unroll 1
                    for (int g = 0; g < kr * 2; ++g) {               // groups of 8 neighbour rows
                        int sir[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) sir[u] = (int)lds32(si_lane + (uint32_t)(g * 32 + u * RPI * 4));
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
                if (lane == 0) { s_kr[s] = kr; mbar_arrive(bar_full + 8 * s); KTC_EV(qs, 0); }
                KTC_T(t3);
                pt[0] += t1 - t0; pt[1] += t2 - t1; pt[2] += t3 - t2;
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
            load_idx(team + NTEAM, b0, b1);
            if (team < nq) {
                if (a0 < Ns) p0 = __ldg(s4 + a0);
                if (a1 < Ns) p1 = __ldg(s4 + a1);
                qx = q[3 * (size_t)(q0 + team)]; qy = q[3 * (size_t)(q0 + team) + 1]; qz = q[3 * (size_t)(q0 + team) + 2];
            }
            for (int qs = team; qs < nq; qs += NTEAM) {
                KTC_T(t0);
                const int n = q0 + qs;
                const int s = qs % NS;
                // current query: relative positions (shadow neighbours sit at x = 3e18: outside every extent)
                const float r0x = a0 < Ns ? p0.x - qx : 3e18f, r0y = p0.y - qy, r0z = p0.z - qz;
                const float r1x = a1 < Ns ? p1.x - qx : 3e18f, r1y = p1.y - qy, r1z = p1.z - qz;
                int nn = (a0 < Ns && p0.w > 0.f ? 1 : 0) + (a1 < Ns && p1.w > 0.f ? 1 : 0);
                // next query: support records and query point now, indices of the one after it
                a0 = b0; a1 = b1;
                p0 = make_float4(0.f, 0.f, 0.f, 0.f); p1 = p0;
                if (qs + NTEAM < nq) {
                    if (a0 < Ns) p0 = __ldg(s4 + a0);
                    if (a1 < Ns) p1 = __ldg(s4 + a1);
                    const size_t nn3 = 3 * (size_t)(n + NTEAM);
                    qx = q[nn3]; qy = q[nn3 + 1]; qz = q[nn3 + 2];
                }
                load_idx(qs + 2 * NTEAM, b0, b1);
                KTC_T(t1);
                if (lane == 0) KTC_WAIT(bar_empty + 8 * s, ((qs / NS) & 1) ^ 1);
                __syncwarp();
                KTC_T(t2);
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
                if (!(dbg & 8)) {
                    uint32_t* brow = reinterpret_cast<uint32_t*>(gen + (size_t)s * Cfg::SLOT + Cfg::A_BYTES + (size_t)k0 * 128) + (lane & 3);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) brow[kk * 32 + (((lane >> 2) ^ kk) << 2)] = hw[kk];   // (k0 + kk) & 7 == kk
                }
                if (tw == 1) nn = __reduce_add_sync(0xffffffffu, nn);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (tw == 1) inv_nn[n] = 1.0f / (float)max(nn, 1);
                    mbar_arrive(bar_full + 8 * s);
                    KTC_EV(qs, tw);
                }
                KTC_T(t3);
                pt[0] += t1 - t0; pt[1] += t2 - t1; pt[2] += t3 - t2;
            }
        }
        if ((dbg & 16) && blockIdx.x == 0 && lane == 0 && prof && team == 0)
            for (int i = 0; i < 3; ++i) prof[tw * 8 + i] = pt[i];
    } else if (warp >= KTC_MMA_WARP) {
        // ================= MMA issuers =================
        const int mw = warp - KTC_MMA_WARP;
        if (lane == 0) {
            // D fp32, A/B fp16, A MN-major (bit 15), B K-major, N = 16, M = 64 / 128
            const uint32_t idesc = (1u << 4) | (1u << 15) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(Cfg::MM >> 4) << 24);
            long long mt[4] = {0, 0, 0, 0};
            for (int qs = 0; qs < nq; ++qs) {
                const int s = qs % NS;
                const int ts = qs / QPT, t = ts % NT;
                if (ts % KTC_MMA_WARPS != mw) continue;
                KTC_T(t0);
                if (dbg & 128) mbar_wait_parked(bar_full + 8 * s, (qs / NS) & 1); else mbar_wait(bar_full + 8 * s, (qs / NS) & 1);
                KTC_T(t1);
                KTC_EV(qs, 3);
                const int kr = (dbg & 2) ? 1 : s_kr[s];
                if (QPT == 1 || (qs & 1) == 0) mbar_wait(bar_dempty + 8 * t, ((ts / NT) & 1) ^ 1);
                tc_fence_after();
                KTC_T(t2);
                KTC_EV(qs, 4);
                const uint32_t a_base = base + s * Cfg::SLOT, b_base = a_base + Cfg::A_BYTES;
                const uint64_t db = make_smem_desc(b_base);
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                    const uint32_t tmem_d = tmem_base + (uint32_t)(t * 16 * MB + mb * 16) + (QPT == 2 ? ((uint32_t)((qs & 1) * 16) << 16) : 0u);
                    for (int ks = 0; ks < kr; ++ks) {
                        const uint64_t da = make_smem_desc_mn(a_base + (uint32_t)mb * 16384u + (uint32_t)ks * 2048u, 8192u, 1024u);
                        tc_mma_f16(tmem_d, da, db + 2 * ks, idesc, ks != 0);
                    }
                }
                tc_commit(bar_empty + 8 * s);
                tc_commit(bar_dfull + 8 * t);
                KTC_T(t3);
                KTC_EV(qs, 5);
                mt[0] += t1 - t0; mt[1] += t2 - t1; mt[2] += t3 - t2;
            }
            if (QPT == 2 && (nq & 1) && ((nq - 1) / 2) % KTC_MMA_WARPS == mw)
                tc_commit(bar_dfull + 8 * (((nq - 1) / 2) % NT));    // the odd last query has no partner
            if ((dbg & 16) && blockIdx.x == 0 && prof && mw == 0)
                for (int i = 0; i < 3; ++i) prof[64 + i] = mt[i];
        }
    } else {
        // ================= epilogue: TMEM -> fp16 -> wf[n, c, k] =================
        const int e = warp;                                          // TMEM lane quarter
        const int nslots = (nq + QPT - 1) / QPT;
        long long et[3] = {0, 0, 0};
        for (int ts = 0; ts < nslots; ++ts) {
            const int t = ts % NT;
            KTC_T(t0);
            if (lane == 0) KTC_WAIT(bar_dfull + 8 * t, (ts / NT) & 1);
            __syncwarp();
            tc_fence_after();
            KTC_T(t1);
            if (e == 0 && lane == 0) KTC_EV(ts * QPT, 6);
            uint32_t v[MB][16];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
                tc_ld16(tmem_base + ((uint32_t)(32 * e) << 16) + (uint32_t)(t * 16 * MB + mb * 16), v[mb]);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_dempty + 8 * t);
            KTC_T(t2);
            int n, c;
            if (QPT == 2) { n = q0 + 2 * ts + (lane >> 4); c = 16 * e + (lane & 15); }
            else { n = q0 + ts; c = 32 * e + lane; }
            if (n < q0 + nq && !(dbg & 4)) {
                // wf[n, c, 0..15]: channel-major, kernel point minor (column K.. are zero: their weights are) — 32 contiguous
                // bytes per lane, 1 KB per warp
                uint4* row = reinterpret_cast<uint4*>(wf + ((size_t)n * CIN + c) * KP_MAX_K);
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                    uint32_t h[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const __half2 t = __floats2half2_rn(__uint_as_float(v[mb][2 * k]), __uint_as_float(v[mb][2 * k + 1]));
                        h[k] = *reinterpret_cast<const uint32_t*>(&t);
                    }
                    row[mb * 128 * 2] = make_uint4(h[0], h[1], h[2], h[3]);
                    row[mb * 128 * 2 + 1] = make_uint4(h[4], h[5], h[6], h[7]);
                }
            }
            KTC_T(t3);
            if (e == 0 && lane == 0) KTC_EV(ts * QPT, 7);
            et[0] += t1 - t0; et[1] += t2 - t1; et[2] += t3 - t2;
        }
        if ((dbg & 16) && blockIdx.x == 0 && lane == 0 && prof)
            for (int i = 0; i < 3; ++i) prof[72 + e * 4 + i] = et[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == KTC_MMA_WARP) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(KTC_TCOLS));
    }
#undef KTC_T
#undef KTC_WAIT
#undef KTC_EV
}

long long* g_ktc_prof = nullptr;   // device buffer of 128 int64 for the in-kernel cycle counters (ktc_dbg bit 16)
int g_ktc_dbg = 0;     // aprb_set_option("ktc_dbg"): diagnostic bits (1 no gather, 2 one K step, 4 no stores, 8 no weights, 16 cycle counters, 32 cp.async.ca)
int g_kpconv_tc = 1;   // aprb_set_option("kpconv_tc"): weighting stage on tcgen05 (fp16 features, Cin in {64, 128, 256}, H <= 64)

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
    if ((g_ktc_dbg & 16) && !g_ktc_prof) { APRB_CUDA_OK(cudaMalloc(&g_ktc_prof, 1024 * sizeof(long long))); }
    if (g_ktc_dbg & 16) APRB_CUDA_OK(cudaMemsetAsync(g_ktc_prof, 0, 1024 * sizeof(long long), st));
    int per = cdiv(Nq, sm_count());
    per = (per + 15) & ~15;                                          // pairs of queries stay in one CTA
    const int grid = cdiv(Nq, per);
    {
        ProfScope ps("kpconv_tc_kernel", st, 1);
        kpconv_tc_kernel<CIN, L1><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(d_q, s4, d_idx, ld, (const __half*)d_x16, d_kp, extent, Nq,
                                                                       Ns, H, K, per, (__half*)d_wf16, d_inv_nn, g_ktc_dbg, g_ktc_prof);
    }
    APRB_LAUNCH_OK();
    if (g_ktc_dbg & 16) {                                            // diagnostic: per-phase cycle counters of CTA 0, team 0
        long long h[1024];
        APRB_CUDA_OK(cudaStreamSynchronize(st));
        APRB_CUDA_OK(cudaMemcpy(h, g_ktc_prof, sizeof(h), cudaMemcpyDeviceToHost));
        const int nq0 = per < Nq ? per : Nq;
        const int nw = (nq0 + Cfg::NTEAM - 1) / Cfg::NTEAM;
        printf("kpconv_tc<%d> CTA0: %d queries, %d teams. team 0 cycles/query [prologue | wait slot | work+arrive]: gather %6.0f %6.0f %6.0f | "
               "weights A %6.0f %6.0f %6.0f | weights B %6.0f %6.0f %6.0f\n", CIN, nq0, Cfg::NTEAM, (double)h[0] / nw, (double)h[1] / nw,
               (double)h[2] / nw, (double)h[8] / nw, (double)h[9] / nw, (double)h[10] / nw, (double)h[16] / nw, (double)h[17] / nw, (double)h[18] / nw);
        printf("  MMA issuer 0 cycles/query [wait full | wait dempty | issue+commit]: %6.0f %6.0f %6.0f\n", (double)h[64] * KTC_MMA_WARPS / nq0,
               (double)h[65] * KTC_MMA_WARPS / nq0, (double)h[66] * KTC_MMA_WARPS / nq0);
        const int nsl = (nq0 + Cfg::QPT - 1) / Cfg::QPT;
        printf("  epilogue warp 0 cycles/slot [wait dfull | ld+arrive | convert+store]: %6.0f %6.0f %6.0f\n", (double)h[72] / nsl, (double)h[73] / nsl,
               (double)h[74] / nsl);
        long long t00 = 0;
        for (int qi = 0; qi < 96; ++qi) for (int e = 0; e < 8; ++e) if (h[128 + qi * 8 + e] && (!t00 || h[128 + qi * 8 + e] < t00)) t00 = h[128 + qi * 8 + e];
        printf("  trace (cycles since first event) q: gather-arrive wA wB | issuer: full-seen slot-ok issued | epilogue: dfull-seen drained\n");
        for (int qi = 0; qi < 96; ++qi) {
            printf("  q%02d:", qi);
            for (int e = 0; e < 8; ++e) printf(" %7lld", h[128 + qi * 8 + e] ? h[128 + qi * 8 + e] - t00 : -1LL);
            printf("\n");
        }
        fflush(stdout);
    }
    return APRB_OK;
}

// Weighted tile wf [Nq, Cin*16] fp16 (channel-major, kernel point minor) and inv_nn [Nq] from fp16 features; s4 = packed support records (x, y, z, flag).
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
