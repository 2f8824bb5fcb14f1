// gemm_tcgen05.cu — TF32 GEMM on Blackwell 5th-gen tensor cores: C[M,N] = (A[M,K] @ Bt[N,K]^T) * rowscale[M].
//
// This is stage C of KPConv (A = kernel-point-weighted features [Nq, K*Cin], Bt = prepared weights [Cout, K*Cin],
// rowscale = 1/neighbor_num; replaces the batched matmul + sum of /root/reference/Predator_APR/models/blocks.py:
// 361-372) and the UnaryBlock nn.Linear (A = features, Bt = mlp.weight [Cout,Cin]; blocks.py:493,500).
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0      : TMA producer — cp.async.bulk.tensor 2D loads of the A (128 x 32 fp32) and B (BN x 32 fp32) k-blocks
//                 into a STAGES-deep ring of 128B-swizzled shared-memory tiles, completion on "full" mbarriers;
//   warp 1      : allocates TMEM (BN fp32 columns) and, through one elected lane, issues tcgen05.mma
//                 (cta_group::1, kind::tf32, M=128, N=BN, K=8; 4 per k-block) with the accumulator in TMEM;
//                 tcgen05.commit releases each smem stage ("empty" mbarriers) and finally signals "tmem_full";
//   warps 2..5  : epilogue — tcgen05.ld (32 lanes x 32 columns per warp and step) -> row scale -> global stores.
// Operands are fp32 in memory; the tensor core reads them as TF32 (producers round to nearest beforehand where the
// 1e-3 parity budget needs it), accumulation is fp32 in TMEM. Out-of-range rows/cols are zero-filled by TMA and masked
// in the epilogue, so M and N need no padding (N % 16 == 0, K % 32 == 0 required).
#include "tc_common.cuh"
#include <cuda_fp16.h>

namespace aprb {

template <int BN>
struct GemmCfg {
    static constexpr int MAX_STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);     // one CTA per SM, deepest ring
    static constexpr int CO_STAGES = BN == 256 ? 2 : (BN == 128 ? 3 : 4);      // two CTAs per SM (<= 112 KB each)
    static constexpr int A_BYTES = GEMM_BM * 128;
    static constexpr int B_BYTES = BN * 128;
    // ring + align slack + barriers. The epilogue's transpose buffer (4 warps x 32 x 36 floats = 18 KB) aliases the
    // A ring: by the time tmem_full fires every TMA write has landed and every MMA that reads the ring has retired.
    static constexpr int smem(int stages) { return (stages < 2 ? 2 : stages) * (A_BYTES + B_BYTES) + 1024 + 256; }
};

// CL = thread-block-cluster size along M: the CL CTAs of a cluster compute CL vertically adjacent 128-row tiles of
// the same BN columns, so they need the SAME B k-blocks: each CTA loads 1/CL of every B tile and TMA-multicasts it to
// all CTAs of the cluster (L2->SM traffic for B divided by CL). A stage is refilled only after all CL CTAs released it
// (the MMA warps commit to the "empty" barriers of every CTA in the cluster).
template <int BN, int CL>
__global__ void __launch_bounds__(192, 2)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                 int kb_per_split, int stages, const float* __restrict__ rowscale, float* __restrict__ C,
                 float* __restrict__ gstat) {
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                    // SWIZZLE_128B tiles need 1024 B alignment
    const int ring = stages < 2 ? 2 : stages;                        // allocated stages (>= 2: room for the epilogue buffer)
    const uint32_t sA = base, sB = base + ring * Cfg::A_BYTES;
    const uint32_t bars = sB + ring * Cfg::B_BYTES;                  // full[stages], empty[stages], tmem_full
    const uint32_t bar_full = bars, bar_empty = bars + 8 * stages, bar_tmem = bars + 16 * stages;
    __shared__ uint32_t s_tmem_base;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * BN;
    // split-K: blockIdx.z owns k-blocks [kb0, kb1); its raw partial tile goes to slice blockIdx.z of C (then C is the
    // [splits, M, N] partial buffer and rowscale is applied by splitk_reduce_kernel)
    const int kb0 = blockIdx.z * kb_per_split, kb1 = min(K / GEMM_BK, kb0 + kb_per_split);
    C += (size_t)blockIdx.z * M * N;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, CL); }
        mbar_init(bar_tmem, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {   // TMEM allocation: BN fp32 accumulator columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                                 // peers' barriers are initialised before any multicast
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1);

    if (warp == 0) {
        if (lane == 0) {                                             // ===== TMA producer =====
            int s = 0; uint32_t ph = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                mbar_arrive_expect_tx(bar_full + 8 * s, Cfg::A_BYTES + Cfg::B_BYTES);
                tma_load_2d(sA + s * Cfg::A_BYTES, &tmA, bar_full + 8 * s, kb * GEMM_BK, m0);
                if (CL == 1) tma_load_2d(sB + s * Cfg::B_BYTES, &tmB, bar_full + 8 * s, kb * GEMM_BK, n0);
                else tma_load_2d_mcast(sB + s * Cfg::B_BYTES + crank * (BN / CL) * 128, &tmB, bar_full + 8 * s, kb * GEMM_BK,
                                       n0 + crank * (BN / CL), kMask);
                if (++s == stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                             // ===== MMA issuer =====
            // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2,
            // B=TF32 [10,13)=2, A/B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GEMM_BM >> 4) << 24);
            int s = 0; uint32_t ph = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(bar_full + 8 * s, ph);
                tc_fence_after();
                const uint64_t da = make_smem_desc(sA + s * Cfg::A_BYTES), db = make_smem_desc(sB + s * Cfg::B_BYTES);
#pragma unroll
                for (int k4 = 0; k4 < GEMM_BK / 8; ++k4)             // +32 bytes (>>4 = 2) per K=8 step inside the atom
                    tc_mma_tf32(tmem_base, da + 2 * k4, db + 2 * k4, idesc, ((kb - kb0) | k4) != 0);
                if (CL == 1) tc_commit(bar_empty + 8 * s);           // frees the stage when these MMAs retire
                else tc_commit_mcast(bar_empty + 8 * s, kMask);      // ... in every CTA of the cluster
                if (++s == stages) { s = 0; ph ^= 1; }
            }
            tc_commit(bar_tmem);                                     // accumulator complete
        }
    } else {                                                         // ===== epilogue (warps 2..5) =====
        mbar_wait(bar_tmem, 0);
        tc_fence_after();
        const int quarter = warp & 3;                                // TMEM lane quarter this warp may access
        const int row_l = quarter * 32 + lane;                       // tile row this thread holds after tcgen05.ld
        const float sc = (rowscale && m0 + row_l < M) ? rowscale[m0 + row_l] : 1.0f;
        // Per-warp 32 x 32 transpose buffer (row stride 36 floats: 16-byte aligned, conflict-free float4 phases), so the
        // global stores are whole 128-byte row segments: lanes 0-7 cover one row's 32 columns, a warp store = 4 rows.
        float* tbuf = reinterpret_cast<float*>(smem_raw + (base - raw)) + (size_t)(warp - 2) * 32 * 36;   // aliases the A ring
        const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(tbuf + lane * 36 + j) = make_float4(__uint_as_float(v[j]) * sc, __uint_as_float(v[j + 1]) * sc,
                                                                              __uint_as_float(v[j + 2]) * sc, __uint_as_float(v[j + 3]) * sc);
            __syncwarp();
            if (n0 + c + sub_c < N) {                                // N % 16 == 0: a float4 is all-in or all-out
#pragma unroll
                for (int r4 = 0; r4 < 32; r4 += 4) {
                    const int r = r4 + sub_r, grow = m0 + quarter * 32 + r;
                    if (grow < M)
                        *reinterpret_cast<float4*>(C + (size_t)grow * N + n0 + c + sub_c) = *reinterpret_cast<const float4*>(tbuf + r * 36 + sub_c);
                }
            }
            // Column statistics of this warp's 32 rows (one group of the normalisation that follows: aprb_instnorm_*_pre):
            // lane = column; mean and M2 = sum of squared deviations from that mean, two passes over the transpose buffer.
            // Only whole groups are recorded; the consumer reads the rows of a ragged last group itself.
            if (gstat && m0 + quarter * 32 + 32 <= M && n0 + c + lane < N) {
                float sum = 0.f;
#pragma unroll
                for (int r = 0; r < 32; ++r) sum += tbuf[r * 36 + lane];
                const float mean = sum * (1.0f / 32.0f);
                float m2 = 0.f;
#pragma unroll
                for (int r = 0; r < 32; ++r) { const float d = tbuf[r * 36 + lane] - mean; m2 = fmaf(d, d, m2); }
                float* gp = gstat + (size_t)((m0 >> 5) + quarter) * 2 * N + n0 + c + lane;
                gp[0] = mean; gp[N] = m2;
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                                 // no CTA exits while a peer may still write to it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN));
    }
}

// ---- persistent variant (no split-K, no cluster) ------------------------------------------------------------------
// One CTA per SM walks output tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (tile t -> column tile t % num_n fastest,
// so CTAs that run together share the A rows in L2). The accumulator is double-buffered in TMEM (2 x BN columns):
// the MMA warp starts tile i+1 in the other buffer while the epilogue warps drain tile i — TMEM -> registers -> smem
// transpose -> HBM, plus the group statistics — so the epilogue (the long part of a small-K Linear) is off the
// critical path, the smem ring keeps streaming across tile boundaries, and barrier init / TMEM allocation / tensor-map
// prefetch are paid once per SM instead of once per tile.
template <int BN>
struct GemmPCfg {
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int A_BYTES = GEMM_BM * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int TBUF = 4 * 32 * 36 * 4;
    static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + TBUF + 1024 + 256;
};

// F16: operands are fp16 (64 elements per 128-byte k-block, kind::f16) instead of TF32-in-fp32 (32 elements).
template <int BN, bool F16>
__global__ void __launch_bounds__(192, 1)
gemm_tf32_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                            int K, int num_n, int total_tiles, const float* __restrict__ rowscale, float* __restrict__ C,
                            float* __restrict__ gstat, const int STAGES) {
    // STAGES (<= Cfg::STAGES) = ring depth of this launch: a shallower ring leaves shared memory for the CTAs of other
    // streams' kernels on the same SM (aprb_set_option("gemm_stages")).
    using Cfg = GemmPCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + STAGES * Cfg::A_BYTES;
    const uint32_t tb = sB + STAGES * Cfg::B_BYTES;                  // epilogue transpose buffers
    const uint32_t bars = tb + Cfg::TBUF;                            // full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]
    const uint32_t bar_full = bars, bar_empty = bars + 8 * STAGES, bar_tfull = bars + 16 * STAGES, bar_tempty = bar_tfull + 16;
    __shared__ uint32_t s_tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int BKE = F16 ? 64 : GEMM_BK;                          // elements per 128-byte k-block
    const int num_kb = K / BKE;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(2 * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == 0) {
        if (lane == 0) {                                             // ===== TMA producer =====
            int s = 0; uint32_t ph = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int m0 = (t / num_n) * GEMM_BM, n0 = (t % num_n) * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    mbar_arrive_expect_tx(bar_full + 8 * s, Cfg::A_BYTES + Cfg::B_BYTES);
                    tma_load_2d(sA + s * Cfg::A_BYTES, &tmA, bar_full + 8 * s, kb * BKE, m0);
                    tma_load_2d(sB + s * Cfg::B_BYTES, &tmB, bar_full + 8 * s, kb * BKE, n0);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                             // ===== MMA issuer =====
            // instruction descriptor: D = F32; A/B format TF32 (2) or F16 (0); K-major; N >> 3, M >> 4
            const uint32_t fmt = F16 ? 0u : 2u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GEMM_BM >> 4) << 24);
            int s = 0; uint32_t ph = 0;
            int i = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
                const int ab = i & 1;
                mbar_wait(bar_tempty + 8 * ab, ((uint32_t)(i >> 1) & 1u) ^ 1u);   // epilogue has drained this buffer
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(ab * BN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(bar_full + 8 * s, ph);
                    tc_fence_after();
                    const uint64_t da = make_smem_desc(sA + s * Cfg::A_BYTES), db = make_smem_desc(sB + s * Cfg::B_BYTES);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {                  // 4 x 32 bytes of K per 128-byte k-block
                        if (F16) tc_mma_f16(acc, da + 2 * k4, db + 2 * k4, idesc, (kb | k4) != 0);
                        else tc_mma_tf32(acc, da + 2 * k4, db + 2 * k4, idesc, (kb | k4) != 0);
                    }
                    tc_commit(bar_empty + 8 * s);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                tc_commit(bar_tfull + 8 * ab);
            }
        }
    } else {                                                         // ===== epilogue (warps 2..5) =====
        const int quarter = warp & 3;
        float* tbuf = reinterpret_cast<float*>(smem_raw + (tb - raw)) + (size_t)(warp - 2) * 32 * 36;
        const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
        int i = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++i) {
            const int ab = i & 1;
            const int m0 = (t / num_n) * GEMM_BM, n0 = (t % num_n) * BN;
            mbar_wait(bar_tfull + 8 * ab, (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
            const int row_l = quarter * 32 + lane;
            const float sc = (rowscale && m0 + row_l < M) ? rowscale[m0 + row_l] : 1.0f;
            const bool whole = m0 + quarter * 32 + 32 <= M;
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * BN + c), v);
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(tbuf + lane * 36 + j) = make_float4(__uint_as_float(v[j]) * sc, __uint_as_float(v[j + 1]) * sc,
                                                                                  __uint_as_float(v[j + 2]) * sc, __uint_as_float(v[j + 3]) * sc);
                __syncwarp();
                if (n0 + c + sub_c < N) {
#pragma unroll
                    for (int r4 = 0; r4 < 32; r4 += 4) {
                        const int r = r4 + sub_r, grow = m0 + quarter * 32 + r;
                        if (grow < M)
                            *reinterpret_cast<float4*>(C + (size_t)grow * N + n0 + c + sub_c) = *reinterpret_cast<const float4*>(tbuf + r * 36 + sub_c);
                    }
                }
                if (gstat && whole && n0 + c + lane < N) {           // group statistics, as in gemm_tf32_kernel
                    float sum = 0.f;
#pragma unroll
                    for (int r = 0; r < 32; ++r) sum += tbuf[r * 36 + lane];
                    const float mean = sum * (1.0f / 32.0f);
                    float m2 = 0.f;
#pragma unroll
                    for (int r = 0; r < 32; ++r) { const float d = tbuf[r * 36 + lane] - mean; m2 = fmaf(d, d, m2); }
                    float* gp = gstat + (size_t)((m0 >> 5) + quarter) * 2 * N + n0 + c + lane;
                    gp[0] = mean; gp[N] = m2;
                }
                __syncwarp();
            }
            tc_fence_before();                                       // all tcgen05.ld of this buffer have completed
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * ab);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN));
    }
}


// ---- recompute path: GEMM whose epilogue IS the normalisation -------------------------------------------------------
// For the small-K Linears that end a ResnetBottleneckBlock (unary2: C/4 -> C, shortcut: C_in -> C; blocks.py:669-681) the
// fp32 output is the largest tensor of the block and exists only to be standardised: written (4C bytes/row), read
// back by the normalisation (4C) and rewritten in fp16 (2C). Recomputing the product is cheaper than that round trip.
// gemm_nrm_f16_kernel runs twice:
//   STATS : the contraction, of which only the per-32-row-group column statistics (mean, M2) leave the SM — plus the
//           rows the normalisation must read from the tensor itself: the ragged last group and the groups that
//           straddle a segment boundary seg_off[1..S-1];
//   apply : the contraction again — bit-identical accumulators: same operands, same K order — finished as
//             y = LeakyReLU( (acc - mean) * rstd  +  shortcut )      shortcut = fp16 residual rows, or, DUAL, a second
//           product (x_sc @ W_sc^T - rmean) * rrstd accumulated by the same CTA in a second TMEM buffer,
//           rounded to the 10-bit mantissa and stored in fp16 (fp32 for the encoder's final output).
// These products have K = 64..1024, so a tile's main loop is short and the kernel lives in its epilogue: 8 epilogue
// warps (two per TMEM lane quarter, 64 of the tile's 128 columns each) instead of the 4 of the general kernel, the
// residual rows of a tile prefetched into registers before the accumulator is awaited, and statistics summed in four
// independent chains. BN = 128; TMEM holds 2 (double buffer) x 1 or 2 accumulators = 256 / 512 columns.
struct NrmArgs {
    int M, N;
    int nkb_main, nkb_sc;            // k-blocks (64 fp16) of the main product and of the shortcut product (0: none)
    int num_n, total_tiles;
    int stages;                      // smem ring depth of this launch
    int park;                        // TMA / MMA lanes wait with the suspend hint instead of polling (aprb_set_option("nrm_park"))
    const int* seg_off;              // S + 1 row offsets of the normalisation segments, or NULL (one segment)
    int S;
    // STATS
    float* gstat;                    // [ceil(M/32)][mean | M2][N]
    float* C;                        // fp32 [M, N]: receives the ragged / straddling groups only
    // apply
    const float* stats;              // [S][nt][mean | rstd][N], nt = 1 + DUAL (tensor 0 = main, 1 = shortcut product)
    const __half* res;               // optional plain fp16 residual [M, N]
    float slope;
    void* out;
    int out16;
};

struct GemmNCfg {
    static constexpr int BN = 128;
    static constexpr int STAGES = 5;
    static constexpr int A_BYTES = GEMM_BM * 128;
    static constexpr int B_BYTES = BN * 128;
    static constexpr int EPI_WARPS = 8;
    static constexpr int TBUF = EPI_WARPS * 32 * 36 * 4;
    static constexpr int SMEM = STAGES * (A_BYTES + B_BYTES) + TBUF + 1024 + 256;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
};

// The kernel lives in its epilogue, so the single TMA and MMA lanes mostly wait: polling, they take a third of the issued
// instructions from the epilogue warps that share their schedulers; parked, the hardware wakes them on the phase flip.
__device__ __forceinline__ void nrm_wait(uint32_t bar, uint32_t parity, int park) {
    if (park) mbar_wait_parked(bar, parity);
    else mbar_wait(bar, parity);
}

__device__ __forceinline__ float nrm_act_round(float v, float slope) {
    v = v >= 0.f ? v : v * slope;
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    return __uint_as_float(u);
}

template <bool DUAL, bool STATS>
__global__ void __launch_bounds__(GemmNCfg::THREADS, 1)
gemm_nrm_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2, const NrmArgs p) {
    using Cfg = GemmNCfg;
    constexpr int BN = Cfg::BN;
    const int STAGES = p.stages;                                     // ring depth of this launch (<= Cfg::STAGES)
    constexpr int ACC = DUAL ? 2 * BN : BN;                          // TMEM columns per tile (main | shortcut)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + STAGES * Cfg::A_BYTES;
    const uint32_t tb = sB + STAGES * Cfg::B_BYTES;
    const uint32_t bars = tb + Cfg::TBUF;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * STAGES, bar_tfull = bars + 16 * STAGES, bar_tempty = bar_tfull + 16;
    __shared__ uint32_t s_tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = p.M, N = p.N;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, Cfg::EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        if (DUAL) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "n"(2 * ACC));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == 0) {
        if (lane == 0) {                                             // ===== TMA producer =====
            int s = 0; uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const int m0 = (t / p.num_n) * GEMM_BM, n0 = (t % p.num_n) * BN;
                if (DUAL)
                    for (int kb = 0; kb < p.nkb_sc; ++kb) {
                        nrm_wait(bar_empty + 8 * s, ph ^ 1, p.park);
                        mbar_arrive_expect_tx(bar_full + 8 * s, Cfg::A_BYTES + Cfg::B_BYTES);
                        tma_load_2d(sA + s * Cfg::A_BYTES, &tmA2, bar_full + 8 * s, kb * 64, m0);
                        tma_load_2d(sB + s * Cfg::B_BYTES, &tmB2, bar_full + 8 * s, kb * 64, n0);
                        if (++s == STAGES) { s = 0; ph ^= 1; }
                    }
                for (int kb = 0; kb < p.nkb_main; ++kb) {
                    nrm_wait(bar_empty + 8 * s, ph ^ 1, p.park);
                    mbar_arrive_expect_tx(bar_full + 8 * s, Cfg::A_BYTES + Cfg::B_BYTES);
                    tma_load_2d(sA + s * Cfg::A_BYTES, &tmA, bar_full + 8 * s, kb * 64, m0);
                    tma_load_2d(sB + s * Cfg::B_BYTES, &tmB, bar_full + 8 * s, kb * 64, n0);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                             // ===== MMA issuer =====
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GEMM_BM >> 4) << 24);   // F16 x F16 -> F32
            int s = 0; uint32_t ph = 0;
            int i = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
                const int ab = i & 1;
                nrm_wait(bar_tempty + 8 * ab, ((uint32_t)(i >> 1) & 1u) ^ 1u, p.park);
                tc_fence_after();
                const uint32_t acc_main = tmem_base + (uint32_t)(ab * ACC), acc_sc = acc_main + BN;
                if (DUAL)
                    for (int kb = 0; kb < p.nkb_sc; ++kb) {
                        nrm_wait(bar_full + 8 * s, ph, p.park);
                        tc_fence_after();
                        const uint64_t da = make_smem_desc(sA + s * Cfg::A_BYTES), db = make_smem_desc(sB + s * Cfg::B_BYTES);
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16(acc_sc, da + 2 * k4, db + 2 * k4, idesc, (kb | k4) != 0);
                        tc_commit(bar_empty + 8 * s);
                        if (++s == STAGES) { s = 0; ph ^= 1; }
                    }
                for (int kb = 0; kb < p.nkb_main; ++kb) {
                    nrm_wait(bar_full + 8 * s, ph, p.park);
                    tc_fence_after();
                    const uint64_t da = make_smem_desc(sA + s * Cfg::A_BYTES), db = make_smem_desc(sB + s * Cfg::B_BYTES);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) tc_mma_f16(acc_main, da + 2 * k4, db + 2 * k4, idesc, (kb | k4) != 0);
                    tc_commit(bar_empty + 8 * s);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                tc_commit(bar_tfull + 8 * ab);
            }
        }
    } else {                                                         // ===== epilogue (warps 2..9) =====
        const int quarter = warp & 3;                                // TMEM lane quarter this warp may read
        const int chalf = (warp - 2) >> 2;                           // which 64 of the tile's 128 columns
        float* tbuf = reinterpret_cast<float*>(smem_raw + (tb - raw)) + (size_t)(warp - 2) * 32 * 36;
        const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
        constexpr int NT = DUAL ? 2 : 1;
        int i = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
            const int ab = i & 1;
            const int m0 = (t / p.num_n) * GEMM_BM, n0 = (t % p.num_n) * BN;
            const int g0 = m0 + quarter * 32;                        // first row of this warp's 32-row group
            const float* stp = nullptr;
            bool store = true, whole = true;
            uint2 rq[2][8];
            if (STATS) {
                whole = g0 + 32 <= M;
                if (whole) {                                         // straddles: some segment starts in (g0, g0 + 31]
                    bool hit = false;
                    for (int s0 = 1; s0 < p.S; s0 += 32) {
                        const int s = s0 + lane;
                        const int o = (s < p.S && p.seg_off) ? p.seg_off[s] : -1;
                        hit |= (o > g0 && o <= g0 + 31);
                    }
                    store = __any_sync(0xffffffffu, hit);
                }
            } else {
                // this thread's accumulator row and the normalisation segment it belongs to
                const int row = min(g0 + lane, M - 1);
                const int seg = (p.seg_off && p.S > 1) ? find_cloud(p.seg_off, p.S, row) : 0;
                stp = p.stats + (size_t)seg * NT * 2 * N + n0;
                if (!DUAL && p.res) {                                // residual rows of the tile: in flight before the wait
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        const int c = chalf * 64 + cc * 32;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int grow = g0 + k * 4 + sub_r;
                            rq[cc][k] = (grow < M && n0 + c + sub_c < N)
                                            ? __ldg(reinterpret_cast<const uint2*>(p.res + (size_t)grow * N + n0 + c + sub_c))
                                            : make_uint2(0u, 0u);
                        }
                    }
                }
            }
            mbar_wait(bar_tfull + 8 * ab, (uint32_t)(i >> 1) & 1u);
            tc_fence_after();
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int c = chalf * 64 + cc * 32;
                uint32_t v[32];
                tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * ACC + c), v);
                if (STATS) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(tbuf + lane * 36 + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                                      __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                } else if (DUAL) {
                    uint32_t w[32];
                    tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * ACC + BN + c), w);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (n0 + c + j < N) {
                            const float4 mean = __ldg(reinterpret_cast<const float4*>(stp + c + j));
                            const float4 rstd = __ldg(reinterpret_cast<const float4*>(stp + N + c + j));
                            const float4 rmean = __ldg(reinterpret_cast<const float4*>(stp + 2 * (size_t)N + c + j));
                            const float4 rrstd = __ldg(reinterpret_cast<const float4*>(stp + 3 * (size_t)N + c + j));
                            o.x = fmaf(__uint_as_float(w[j]) - rmean.x, rrstd.x, (__uint_as_float(v[j]) - mean.x) * rstd.x);
                            o.y = fmaf(__uint_as_float(w[j + 1]) - rmean.y, rrstd.y, (__uint_as_float(v[j + 1]) - mean.y) * rstd.y);
                            o.z = fmaf(__uint_as_float(w[j + 2]) - rmean.z, rrstd.z, (__uint_as_float(v[j + 2]) - mean.z) * rstd.z);
                            o.w = fmaf(__uint_as_float(w[j + 3]) - rmean.w, rrstd.w, (__uint_as_float(v[j + 3]) - mean.w) * rstd.w);
                        }
                        *reinterpret_cast<float4*>(tbuf + lane * 36 + j) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (n0 + c + j < N) {
                            const float4 mean = __ldg(reinterpret_cast<const float4*>(stp + c + j));
                            const float4 rstd = __ldg(reinterpret_cast<const float4*>(stp + N + c + j));
                            o.x = (__uint_as_float(v[j]) - mean.x) * rstd.x;
                            o.y = (__uint_as_float(v[j + 1]) - mean.y) * rstd.y;
                            o.z = (__uint_as_float(v[j + 2]) - mean.z) * rstd.z;
                            o.w = (__uint_as_float(v[j + 3]) - mean.w) * rstd.w;
                        }
                        *reinterpret_cast<float4*>(tbuf + lane * 36 + j) = o;
                    }
                }
                __syncwarp();
                if (STATS) {
                    if (store && n0 + c + sub_c < N) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int r = k * 4 + sub_r, grow = g0 + r;
                            if (grow < M)
                                *reinterpret_cast<float4*>(p.C + (size_t)grow * N + n0 + c + sub_c) = *reinterpret_cast<const float4*>(tbuf + r * 36 + sub_c);
                        }
                    }
                    if (whole && n0 + c + lane < N) {                // lane = column: four independent chains over the rows
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                        for (int r = 0; r < 32; r += 4) {
                            s0 += tbuf[r * 36 + lane]; s1 += tbuf[(r + 1) * 36 + lane];
                            s2 += tbuf[(r + 2) * 36 + lane]; s3 += tbuf[(r + 3) * 36 + lane];
                        }
                        const float mean = ((s0 + s1) + (s2 + s3)) * (1.0f / 32.0f);
                        float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
                        for (int r = 0; r < 32; r += 4) {
                            const float d0 = tbuf[r * 36 + lane] - mean, d1 = tbuf[(r + 1) * 36 + lane] - mean;
                            const float d2 = tbuf[(r + 2) * 36 + lane] - mean, d3 = tbuf[(r + 3) * 36 + lane] - mean;
                            q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
                        }
                        float* gp = p.gstat + (size_t)(g0 >> 5) * 2 * N + n0 + c + lane;
                        gp[0] = mean; gp[N] = (q0 + q1) + (q2 + q3);
                    }
                } else if (n0 + c + sub_c < N) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int r = k * 4 + sub_r, grow = g0 + r;
                        if (grow < M) {
                            float4 o = *reinterpret_cast<const float4*>(tbuf + r * 36 + sub_c);
                            const size_t e = (size_t)grow * N + n0 + c + sub_c;
                            if (!DUAL && p.res) {
                                const float2 q0 = __half22float2(*reinterpret_cast<const __half2*>(&rq[cc][k].x));
                                const float2 q1 = __half22float2(*reinterpret_cast<const __half2*>(&rq[cc][k].y));
                                o.x += q0.x; o.y += q0.y; o.z += q1.x; o.w += q1.y;
                            }
                            o.x = nrm_act_round(o.x, p.slope); o.y = nrm_act_round(o.y, p.slope);
                            o.z = nrm_act_round(o.z, p.slope); o.w = nrm_act_round(o.w, p.slope);
                            if (p.out16) {
                                uint2 u;
                                u.x = pack_half2_sat(o.x, o.y); u.y = pack_half2_sat(o.z, o.w);
                                *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + e) = u;
                            } else {
                                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + e) = o;
                            }
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * ab);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * ACC));
    }
}

// C[m,n] = (sum_s part[s,m,n]) * rowscale[m]; fixed summation order (deterministic split-K)
__global__ void splitk_reduce_kernel(const float4* __restrict__ part, int splits, size_t mn4, int n4,
                                     const float* __restrict__ rowscale, float4* __restrict__ C) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mn4) return;
    float4 a = part[i];
    for (int s = 1; s < splits; ++s) {
        const float4 b = part[(size_t)s * mn4 + i];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    const float sc = rowscale ? rowscale[i / n4] : 1.0f;
    C[i] = make_float4(a.x * sc, a.y * sc, a.z * sc, a.w * sc);
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            (void)cudaGetLastError();
    }
    return fn;
}

// 2-D fp32 row-major [rows, cols] tensor, box = [box_rows, 32 cols], 128B swizzle, zero OOB fill
int make_tmap(CUtensorMap* tm, const float* ptr, int rows, int cols, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return APRB_ERR_CUDA; }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)cols * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d box_rows=%d", (int)r, rows, cols, box_rows); return APRB_ERR_CUDA; }
    return APRB_OK;
}

int make_tmap_f16(CUtensorMap* tm, const void* ptr, int rows, int cols, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return APRB_ERR_CUDA; }
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)ptr, gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(f16) failed (%d) rows=%d cols=%d box_rows=%d", (int)r, rows, cols, box_rows); return APRB_ERR_CUDA; }
    return APRB_OK;
}

bool gemm_tf32_supported(int M, int N, int K) {
    return M >= 1 && N >= 16 && N % 16 == 0 && K >= GEMM_BK && K % GEMM_BK == 0;
}

int g_gemm_costages = 1;  // aprb_set_option("gemm_costages"): shallow rings / several CTAs per SM when the grid allows

// Timer label of the next GEMM launches of this thread: the KPConv contraction tags itself ("kpconv_gemm_kernel") so that the
// bench line can report the whole KPConv operator (weighting + contraction) apart from the Linear layers.
static thread_local const char* t_gemm_label = nullptr;
void set_gemm_label(const char* label) { t_gemm_label = label; }
static inline const char* gemm_label() { return t_gemm_label ? t_gemm_label : "gemm_tf32_kernel"; }

template <int BN, int CL>
static int launch_gemm(const float* A, const float* Bt, int M, int N, int K, int splits, int kb_per_split,
                       const float* rowscale, float* C, float* gstat, cudaStream_t st) {
    // Ring depth. With at least two tiles per SM in the grid, a shallower ring lets two (or more) CTAs share an SM so
    // one tile's epilogue (TMEM -> registers -> HBM, the long part of a small-K Linear) overlaps its neighbour's main
    // loop; otherwise one CTA per SM keeps the deepest ring.
    const int tiles = cdiv(N, BN) * cdiv(M, GEMM_BM) * splits;
    int stages = GemmCfg<BN>::MAX_STAGES;
    // (measured, tools/gemm_bench.py: the 2-stage ring of BN = 256 starves tensor-bound contractions with K >= 1024)
    if (g_gemm_costages && tiles >= 2 * sm_count() && (BN < 256 || K <= 512)) stages = GemmCfg<BN>::CO_STAGES;
    stages = min(stages, max(1, kb_per_split));
    const int smem = GemmCfg<BN>::smem(stages);
    CUtensorMap tmA, tmB;
    int rc = make_tmap(&tmA, A, M, K, GEMM_BM);
    if (rc) return rc;
    rc = make_tmap(&tmB, Bt, N, K, BN / CL);                        // each CTA loads a BN/CL-row slice of the B tile
    if (rc) return rc;
    static bool attr_set[64] = {};                              // function attributes are per device
    const int dev_i = current_device() & 63;
    if (!attr_set[dev_i]) {
        APRB_CUDA_OK(cudaFuncSetAttribute(gemm_tf32_kernel<BN, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          GemmCfg<BN>::smem(GemmCfg<BN>::MAX_STAGES)));
        attr_set[dev_i] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cdiv(N, BN), cdiv(cdiv(M, GEMM_BM), CL) * CL, splits);   // grid.y padded to whole clusters
    cfg.blockDim = dim3(192);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = CL; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    {
        ProfScope ps(gemm_label(), st, 1);
        APRB_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_tf32_kernel<BN, CL>, tmA, tmB, M, N, K, kb_per_split, stages, rowscale, C, gstat));
    }
    return APRB_OK;
}

int g_nrm_ctas = 1;          // aprb_set_option("nrm_ctas"): 2 = two CTAs per SM for the statistics pass (measured: 0.46 -> 0.41 ms per call alone, clouds/s unchanged)
int g_nrm_park = 0;          // aprb_set_option("nrm_park"): parked mbarrier waits for the TMA / MMA lanes of gemm_nrm_f16_kernel (measured: no gain)
int g_gemm_stages = 0;       // aprb_set_option("gemm_stages"): cap on the smem ring depth of the persistent kernels (0 = deepest)
int g_gemm_bn = 0;           // aprb_set_option("gemm_bn"): force the persistent kernel's tile width (0 = by wave count)
int g_gemm_persistent = 1;   // aprb_set_option("gemm_persistent"): persistent double-buffered kernel when no split-K is needed

template <int BN, bool F16>
static int launch_gemm_persistent(const void* A, const void* Bt, int M, int N, int K, const float* rowscale, float* C,
                                  float* gstat, cudaStream_t st) {
    CUtensorMap tmA, tmB;
    int rc = F16 ? make_tmap_f16(&tmA, A, M, K, GEMM_BM) : make_tmap(&tmA, (const float*)A, M, K, GEMM_BM);
    if (rc) return rc;
    rc = F16 ? make_tmap_f16(&tmB, Bt, N, K, BN) : make_tmap(&tmB, (const float*)Bt, N, K, BN);
    if (rc) return rc;
    static bool attr_set[64] = {};                              // function attributes are per device
    const int dev_i = current_device() & 63;
    if (!attr_set[dev_i]) {
        APRB_CUDA_OK(cudaFuncSetAttribute(gemm_tf32_persistent_kernel<BN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmPCfg<BN>::SMEM));
        attr_set[dev_i] = true;
    }
    const int num_n = cdiv(N, BN), total = num_n * cdiv(M, GEMM_BM);
    const int grid = min(total, sm_count());
    const int stages = g_gemm_stages >= 2 ? min(g_gemm_stages, GemmPCfg<BN>::STAGES) : GemmPCfg<BN>::STAGES;
    const int smem = GemmPCfg<BN>::SMEM - (GemmPCfg<BN>::STAGES - stages) * (GemmPCfg<BN>::A_BYTES + GemmPCfg<BN>::B_BYTES);
    {
        ProfScope ps(gemm_label(), st, 1);
        gemm_tf32_persistent_kernel<BN, F16><<<grid, 192, smem, st>>>(tmA, tmB, M, N, K, num_n, total, rowscale, C, gstat, stages);
    }
    APRB_LAUNCH_OK();
    return APRB_OK;
}

extern int g_kpconv_chunk_mb;
extern int g_kpw_version;
extern int g_kpw_fh;
extern int g_fuse_stats;
extern int g_kpconv_f16;
extern int g_act_f16;
extern int g_dbg_skip_d2h;
extern int g_host_zero_copy;
extern int g_kpconv_fused;
extern int g_kpconv_tc;
extern int g_ktc_dbg;
extern int g_gemm_apply;
extern int g_blocking_sync;
int g_gemm_cluster = 1;   // 1 disables the cluster/multicast path (aprb_set_option)

size_t gemm_tf32_ws_bytes(int M, int N) {   // split-K partial tiles (up to 8 splits), only when split-K can trigger
    const int bn = N >= 256 ? 256 : (N > 64 ? 128 : 64);
    if (cdiv(M, GEMM_BM) * cdiv(N, bn) >= sm_count()) return 256;
    return align256((size_t)8 * M * N * sizeof(float)) + 256;
}

// d_ws may be NULL (no split-K). Tile width: the widest BN (fewest re-reads of A); split-K fills the SMs when the
// output grid alone cannot.
// d_gstat (optional): per 32-row group column statistics of C, [ceil(M/32)][mean | M2][N], written by the epilogue when the
// product is finished there (no split-K); *stats_written tells the caller whether it was.
int gemm_tf32_rowscale(const float* d_A, const float* d_Bt, int M, int N, int K, const float* d_rowscale, float* d_C,
                       void* d_ws, size_t ws_bytes, cudaStream_t st, float* d_gstat, int* stats_written) {
    if (stats_written) *stats_written = 0;
    if (!gemm_tf32_supported(M, N, K)) { set_error("gemm_tf32: unsupported shape M=%d N=%d K=%d", M, N, K); return APRB_ERR_UNSUPPORTED; }
    if (((uintptr_t)d_A | (uintptr_t)d_Bt | (uintptr_t)d_C) & 15) { set_error("gemm_tf32: operands must be 16-byte aligned"); return APRB_ERR_INVALID; }
    const int mt = cdiv(M, GEMM_BM), sms = sm_count(), num_kb = K / GEMM_BK;
    int bn = N >= 256 ? 256 : (N > 64 ? 128 : 64);
    int splits = 1;
    const int tiles = mt * cdiv(N, bn);
    if (d_ws && tiles < sms) {
        splits = min(min(sms / tiles, 8), max(1, num_kb / 8));       // >= 8 k-blocks per split
        while (splits > 1 && (size_t)splits * M * N * sizeof(float) > ws_bytes) --splits;
    }
    if (splits <= 1 && tiles < sms / 2) {                            // no split-K possible: fall back to narrower tiles
        while (bn > 64 && mt * cdiv(N, bn) < sms) bn >>= 1;
    }
    if (splits <= 1 && g_gemm_persistent) {
        // tile width by wave count: cost = ceil(tiles / SMs) * BN / eff(BN) (narrow tiles re-read A and run the tensor
        // pipe less efficiently, but quantise better on 148 SMs)
        int best = 0; double best_cost = 1e30;
        const int cand[3] = {256, 128, 64};
        const double eff[3] = {1.0, 0.95, 0.80};
        for (int c = 0; c < 3; ++c) {
            if (cand[c] > 64 && N < cand[c]) continue;
            if (g_gemm_bn && cand[c] != g_gemm_bn && !(g_gemm_bn > N && cand[c] == 64)) continue;
            const double cost = (double)cdiv(mt * cdiv(N, cand[c]), sms) * cand[c] / eff[c];
            if (cost < best_cost) { best_cost = cost; best = cand[c]; }
        }
        if (stats_written && d_gstat) *stats_written = 1;
        if (best == 256) return launch_gemm_persistent<256, false>(d_A, d_Bt, M, N, K, d_rowscale, d_C, d_gstat, st);
        if (best == 128) return launch_gemm_persistent<128, false>(d_A, d_Bt, M, N, K, d_rowscale, d_C, d_gstat, st);
        return launch_gemm_persistent<64, false>(d_A, d_Bt, M, N, K, d_rowscale, d_C, d_gstat, st);
    }
    int kps = cdiv(num_kb, splits);
    splits = cdiv(num_kb, kps);
    float* out = splits > 1 ? (float*)d_ws : d_C;
    const float* rs = splits > 1 ? nullptr : d_rowscale;
    float* gs = splits > 1 ? nullptr : d_gstat;
    if (gs && stats_written) *stats_written = 1;
    int rc;
    const int cl = (g_gemm_cluster >= 2 && mt >= 2) ? 2 : 1;       // B-tile multicast across 2 vertically adjacent tiles
    if (bn == 256) rc = cl == 2 ? launch_gemm<256, 2>(d_A, d_Bt, M, N, K, splits, kps, rs, out, gs, st) : launch_gemm<256, 1>(d_A, d_Bt, M, N, K, splits, kps, rs, out, gs, st);
    else if (bn == 128) rc = cl == 2 ? launch_gemm<128, 2>(d_A, d_Bt, M, N, K, splits, kps, rs, out, gs, st) : launch_gemm<128, 1>(d_A, d_Bt, M, N, K, splits, kps, rs, out, gs, st);
    else rc = cl == 2 ? launch_gemm<64, 2>(d_A, d_Bt, M, N, K, splits, kps, rs, out, gs, st) : launch_gemm<64, 1>(d_A, d_Bt, M, N, K, splits, kps, rs, out, gs, st);
    if (rc || splits == 1) return rc;
    const size_t mn4 = (size_t)M * N / 4;
    APRB_TIMED("splitk_reduce_kernel", st, 1, (splitk_reduce_kernel<<<cdiv((long long)mn4, 256), 256, 0, st>>>(
        (const float4*)d_ws, splits, mn4, N / 4, d_rowscale, (float4*)d_C)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

// C[M,N] (fp32) = (A[M,K] @ Bt[N,K]^T) * rowscale[M] with fp16 operands (fp32 accumulation in TMEM): the KPConv
// contraction when the weighted tile is stored in fp16. Persistent kernel only (no split-K). K % 64 == 0, N % 16 == 0.
bool gemm_f16_supported(int M, int N, int K) { return M >= 1 && N >= 16 && N % 16 == 0 && K >= 64 && K % 64 == 0; }

int gemm_f16_rowscale(const void* d_A, const void* d_Bt, int M, int N, int K, const float* d_rowscale, float* d_C,
                      cudaStream_t st, float* d_gstat, int* stats_written) {
    if (stats_written) *stats_written = 0;
    if (!gemm_f16_supported(M, N, K)) { set_error("gemm_f16: unsupported shape M=%d N=%d K=%d", M, N, K); return APRB_ERR_UNSUPPORTED; }
    if (((uintptr_t)d_A | (uintptr_t)d_Bt | (uintptr_t)d_C) & 15) { set_error("gemm_f16: operands must be 16-byte aligned"); return APRB_ERR_INVALID; }
    const int mt = cdiv(M, GEMM_BM), sms = sm_count();
    int best = 0; double best_cost = 1e30;
    const int cand[3] = {256, 128, 64};
    const double eff[3] = {1.0, 0.95, 0.80};
    for (int c = 0; c < 3; ++c) {
        if (cand[c] > 64 && N < cand[c]) continue;
        const double cost = (double)cdiv(mt * cdiv(N, cand[c]), sms) * cand[c] / eff[c];
        if (cost < best_cost) { best_cost = cost; best = cand[c]; }
    }
    if (stats_written && d_gstat) *stats_written = 1;
    if (best == 256) return launch_gemm_persistent<256, true>(d_A, d_Bt, M, N, K, d_rowscale, d_C, d_gstat, st);
    if (best == 128) return launch_gemm_persistent<128, true>(d_A, d_Bt, M, N, K, d_rowscale, d_C, d_gstat, st);
    return launch_gemm_persistent<64, true>(d_A, d_Bt, M, N, K, d_rowscale, d_C, d_gstat, st);
}

template <bool DUAL, bool STATS>
static int launch_gemm_nrm(const void* A, const void* Bt, int K, const void* A2, const void* Bt2, int K2, NrmArgs p,
                           cudaStream_t st) {
    constexpr int BN = GemmNCfg::BN;
    CUtensorMap tmA, tmB, tmA2, tmB2;
    int rc = make_tmap_f16(&tmA, A, p.M, K, GEMM_BM);
    if (rc) return rc;
    rc = make_tmap_f16(&tmB, Bt, p.N, K, BN);
    if (rc) return rc;
    tmA2 = tmA; tmB2 = tmB;
    if (DUAL) {
        rc = make_tmap_f16(&tmA2, A2, p.M, K2, GEMM_BM);
        if (rc) return rc;
        rc = make_tmap_f16(&tmB2, Bt2, p.N, K2, BN);
        if (rc) return rc;
    }
    static bool attr_set[64] = {};                              // function attributes are per device
    const int dev_i = current_device() & 63;
    if (!attr_set[dev_i]) {
        APRB_CUDA_OK(cudaFuncSetAttribute(gemm_nrm_f16_kernel<DUAL, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmNCfg::SMEM));
        attr_set[dev_i] = true;
    }
    p.nkb_main = K / 64; p.nkb_sc = DUAL ? K2 / 64 : 0;
    p.num_n = cdiv(p.N, BN); p.total_tiles = p.num_n * cdiv(p.M, GEMM_BM);
    p.stages = g_gemm_stages >= 2 ? min(g_gemm_stages, GemmNCfg::STAGES) : GemmNCfg::STAGES;
    // aprb_set_option("nrm_ctas", 2): two CTAs per SM for the single-product forms (256 of the 512 TMEM columns each, a 2-stage
    // ring each): the kernel is bound by its per-tile epilogue latency chain, a second CTA overlaps it with another tile's
    const bool two = g_nrm_ctas == 2 && !DUAL && STATS;              // the apply forms need 156 registers x 320 threads: one CTA per SM
    if (two) p.stages = 2;
    p.park = g_nrm_park;
    const int smem = GemmNCfg::SMEM - (GemmNCfg::STAGES - p.stages) * (GemmNCfg::A_BYTES + GemmNCfg::B_BYTES);
    const int grid = min(p.total_tiles, (two ? 2 : 1) * sm_count());
    {
        ProfScope ps(STATS ? "gemm_nrm_stats_kernel" : "gemm_nrm_apply_kernel", st, 1);
        gemm_nrm_f16_kernel<DUAL, STATS><<<grid, GemmNCfg::THREADS, smem, st>>>(tmA, tmB, tmA2, tmB2, p);
    }
    APRB_LAUNCH_OK();
    return APRB_OK;
}

// Statistics pass of the recompute path: group statistics of A @ Bt^T; d_C receives the ragged / straddling groups only.
int gemm_stats_f16(const void* d_A, const void* d_Bt, int M, int N, int K, const int* d_seg_off, int S, float* d_gstat,
                   float* d_C, cudaStream_t st) {
    if (!gemm_f16_supported(M, N, K) || N % 4 != 0) { set_error("gemm_stats_f16: unsupported shape M=%d N=%d K=%d", M, N, K); return APRB_ERR_UNSUPPORTED; }
    if (((uintptr_t)d_A | (uintptr_t)d_Bt | (uintptr_t)d_C | (uintptr_t)d_gstat) & 15) { set_error("gemm_stats_f16: operands must be 16-byte aligned"); return APRB_ERR_INVALID; }
    NrmArgs p = {};
    p.M = M; p.N = N; p.seg_off = d_seg_off; p.S = S; p.gstat = d_gstat; p.C = d_C;
    return launch_gemm_nrm<false, true>(d_A, d_Bt, K, nullptr, nullptr, 0, p, st);
}

// y = LeakyReLU((A @ Bt^T - mean) * rstd + shortcut), see gemm_nrm_f16_kernel. A2 / Bt2 (optional): the shortcut is a
// second product standardised with stats tensor 1; res16 (optional, exclusive with A2): plain fp16 residual rows.
int gemm_apply_f16(const void* d_A, const void* d_Bt, int M, int N, int K, const void* d_A2, const void* d_Bt2, int K2,
                   const int* d_seg_off, int S, const float* d_stats, const void* d_res16, float slope, void* d_out,
                   int out16, cudaStream_t st) {
    if (!gemm_f16_supported(M, N, K) || N % 4 != 0 || (d_A2 && (K2 < 64 || K2 % 64 != 0))) {
        set_error("gemm_apply_f16: unsupported shape M=%d N=%d K=%d K2=%d", M, N, K, K2);
        return APRB_ERR_UNSUPPORTED;
    }
    if (((uintptr_t)d_A | (uintptr_t)d_Bt | (uintptr_t)d_out | (uintptr_t)d_stats | (uintptr_t)(d_A2 ? d_A2 : d_A) |
         (uintptr_t)(d_Bt2 ? d_Bt2 : d_Bt) | (uintptr_t)(d_res16 ? d_res16 : d_A)) & 15) {
        set_error("gemm_apply_f16: operands must be 16-byte aligned");
        return APRB_ERR_INVALID;
    }
    NrmArgs p = {};
    p.M = M; p.N = N; p.seg_off = d_seg_off; p.S = S; p.stats = d_stats; p.res = (const __half*)d_res16; p.slope = slope;
    p.out = d_out; p.out16 = out16;
    if (d_A2) return launch_gemm_nrm<true, false>(d_A, d_Bt, K, d_A2, d_Bt2, K2, p, st);
    return launch_gemm_nrm<false, false>(d_A, d_Bt, K, nullptr, nullptr, 0, p, st);
}

}  // namespace aprb

extern "C" int aprb_set_option(const char* name, int value) {
    using namespace aprb;
    APRB_REQUIRE(name, "null option name");
    if (strcmp(name, "gemm_cluster") == 0) { g_gemm_cluster = value; return APRB_OK; }
    if (strcmp(name, "gemm_costages") == 0) { g_gemm_costages = value; return APRB_OK; }
    if (strcmp(name, "gemm_persistent") == 0) { g_gemm_persistent = value; return APRB_OK; }
    if (strcmp(name, "gemm_bn") == 0) { g_gemm_bn = value; return APRB_OK; }
    if (strcmp(name, "gemm_stages") == 0) { g_gemm_stages = value; return APRB_OK; }
    if (strcmp(name, "nrm_park") == 0) { g_nrm_park = value; return APRB_OK; }
    if (strcmp(name, "nrm_ctas") == 0) { g_nrm_ctas = value; return APRB_OK; }
    if (strcmp(name, "kpconv_chunk_mb") == 0) { g_kpconv_chunk_mb = value; return APRB_OK; }
    if (strcmp(name, "kpw_version") == 0) { g_kpw_version = value; return APRB_OK; }
    if (strcmp(name, "kpw_fh") == 0) { g_kpw_fh = value; return APRB_OK; }
    if (strcmp(name, "fuse_stats") == 0) { g_fuse_stats = value; return APRB_OK; }
    if (strcmp(name, "kpconv_f16") == 0) { g_kpconv_f16 = value; return APRB_OK; }
    if (strcmp(name, "act_f16") == 0) { g_act_f16 = value; return APRB_OK; }
    if (strcmp(name, "dbg_skip_d2h") == 0) { g_dbg_skip_d2h = value; return APRB_OK; }
    if (strcmp(name, "host_zero_copy") == 0) { g_host_zero_copy = value; return APRB_OK; }
    if (strcmp(name, "kpconv_fused") == 0) { g_kpconv_fused = value; return APRB_OK; }
    if (strcmp(name, "kpconv_tc") == 0) { g_kpconv_tc = value; return APRB_OK; }
    if (strcmp(name, "ktc_dbg") == 0) { g_ktc_dbg = value; return APRB_OK; }
    if (strcmp(name, "gemm_apply") == 0) { g_gemm_apply = value; return APRB_OK; }
    if (strcmp(name, "blocking_sync") == 0) { g_blocking_sync = value; return APRB_OK; }
    set_error("aprb_set_option: unknown option %s", name);
    return APRB_ERR_INVALID;
}

extern "C" size_t aprb_linear_tf32_ws_bytes(int N, int Cin, int Cout) {
    (void)Cin;
    if (N < 0 || Cout < 0) return 0;
    return aprb::gemm_tf32_ws_bytes(N > 0 ? N : 1, Cout > 0 ? Cout : 1);
}

extern "C" int aprb_linear_tf32(const float* d_x, const float* d_W, int N, int Cin, int Cout, float* d_y, void* d_ws,
                                size_t ws_bytes, void* stream) {
    return aprb_linear_tf32_stats(d_x, d_W, N, Cin, Cout, d_y, nullptr, nullptr, d_ws, ws_bytes, stream);
}

extern "C" int aprb_linear_f16_stats(const void* d_x16, const void* d_W16, int N, int Cin, int Cout, float* d_y,
                                     float* d_gstat, int* stats_written, void* stream) {
    using namespace aprb;
    if (stats_written) *stats_written = 0;
    APRB_REQUIRE(N >= 0 && Cin >= 1 && Cout >= 1, "bad shape");
    APRB_REQUIRE(!d_gstat || stats_written, "stats_written must be given with d_gstat");
    if (N == 0) return APRB_OK;
    APRB_REQUIRE(d_x16 && d_W16 && d_y, "null pointer");
    return gemm_f16_rowscale(d_x16, d_W16, N, Cout, Cin, nullptr, d_y, (cudaStream_t)stream, d_gstat, stats_written);
}

extern "C" int aprb_linear_f16_stats_ragged(const void* d_x16, const void* d_W16, int N, int Cin, int Cout, float* d_y,
                                            float* d_gstat, const int32_t* d_seg_off, int S, void* stream) {
    using namespace aprb;
    APRB_REQUIRE(N >= 0 && Cin >= 1 && Cout >= 1 && S >= 1, "bad shape");
    APRB_REQUIRE(S == 1 || d_seg_off, "segment offsets required when S > 1");
    if (N == 0) return APRB_OK;
    APRB_REQUIRE(d_x16 && d_W16 && d_y && d_gstat, "null pointer");
    return gemm_stats_f16(d_x16, d_W16, N, Cout, Cin, d_seg_off, S, d_gstat, d_y, (cudaStream_t)stream);
}

extern "C" int aprb_linear_f16_norm_apply(const void* d_x16, const void* d_W16, int N, int Cin, int Cout,
                                          const void* d_sc16, const void* d_Wsc16, int Csc, const void* d_res16,
                                          const int32_t* d_seg_off, int S, const float* d_stats, float slope, void* d_y,
                                          int out_is_f16, void* stream) {
    using namespace aprb;
    APRB_REQUIRE(N >= 0 && Cin >= 1 && Cout >= 1 && S >= 1, "bad shape");
    APRB_REQUIRE(S == 1 || d_seg_off, "segment offsets required when S > 1");
    APRB_REQUIRE((d_sc16 == nullptr) == (d_Wsc16 == nullptr), "shortcut operand and weights go together");
    APRB_REQUIRE(!(d_sc16 && d_res16), "either a shortcut product or a plain residual, not both");
    if (N == 0) return APRB_OK;
    APRB_REQUIRE(d_x16 && d_W16 && d_y && d_stats, "null pointer");
    return gemm_apply_f16(d_x16, d_W16, N, Cout, Cin, d_sc16, d_Wsc16, Csc, d_seg_off, S, d_stats, d_res16, slope, d_y,
                          out_is_f16, (cudaStream_t)stream);
}

extern "C" size_t aprb_group_stats_bytes(int N, int C) {
    if (N < 0 || C < 1) return 0;
    return aprb::align256((size_t)aprb::cdiv(N > 0 ? N : 1, 32) * 2 * C * sizeof(float));
}

extern "C" int aprb_linear_tf32_stats(const float* d_x, const float* d_W, int N, int Cin, int Cout, float* d_y,
                                      float* d_gstat, int* stats_written, void* d_ws, size_t ws_bytes, void* stream) {
    using namespace aprb;
    if (stats_written) *stats_written = 0;
    APRB_REQUIRE(N >= 0 && Cin >= 1 && Cout >= 1, "bad shape");
    APRB_REQUIRE(!d_gstat || stats_written, "stats_written must be given with d_gstat");
    if (N == 0) return APRB_OK;
    APRB_REQUIRE(d_x && d_W && d_y, "null pointer");
    return gemm_tf32_rowscale(d_x, d_W, N, Cout, Cin, nullptr, d_y, d_ws, ws_bytes, (cudaStream_t)stream, d_gstat, stats_written);
}
