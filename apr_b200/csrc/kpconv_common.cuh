// kpconv_common.cuh — device helpers shared by the KPConv kernels (kpconv.cu: stage A+B producers; kpconv_fused.cu:
// the fused gather -> influence -> tcgen05 contraction kernel): per-row neighbour geometry and the linear influence
// max(0, 1 - d / KP_extent) of /root/reference/Predator_APR/models/blocks.py:269-289, :328-329.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace aprb {

constexpr int KP_MAX_K = 16;

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float sqrt_approx(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
// Value whose TRUNCATION to TF32 (what the tensor core does with an fp32 operand) equals round-to-nearest (ties away
// from zero in magnitude) of t: one integer add instead of the 3-instruction cvt.rna emulation. Exact for finite t;
// +-inf becomes NaN (a feature table that holds inf is garbage either way).
__device__ __forceinline__ float pre_round_tf32(float t) { return __uint_as_float(__float_as_uint(t) + 0x1000u); }
__device__ __forceinline__ float round_tf32(float t) {
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(t));
    return __uint_as_float(u);
}

template <int NH>
struct RowGeom {          // lanes = neighbours: element offset of the support's feature row and position relative to the query
    int sio[NH];          // byte offset of the support's feature row, si * row_bytes (shadow neighbours: 0, never selected)
    float rx[NH], ry[NH], rz[NH];   // shadow neighbours sit at x = 3e18: outside every extent
    unsigned any[NH];               // warp ballot: does this group of 32 neighbours hold a valid one (warp-uniform)
};

template <typename IdxT, int NH>
__device__ __forceinline__ int load_row_geom(const float* __restrict__ q, const float4* __restrict__ s4,
                                             const IdxT* __restrict__ idx, int ld, int n, int Ns, int H, int row_bytes, int lane,
                                             RowGeom<NH>& g) {
    const float qx = q[3 * (size_t)n], qy = q[3 * (size_t)n + 1], qz = q[3 * (size_t)n + 2];
    int nn = 0;
#pragma unroll
    for (int j = 0; j < NH; ++j) {
        const int h = j * 32 + lane;
        int si = Ns;
        if (h < H) {
            const long long v = (long long)idx[(size_t)n * ld + h];
            si = (v >= 0 && v < Ns) ? (int)v : Ns;
        }
        g.sio[j] = 0;
        g.rx[j] = 3e18f; g.ry[j] = 0.f; g.rz[j] = 0.f;
        if (si < Ns) {
            const float4 p = __ldg(s4 + si);
            nn += p.w > 0.f ? 1 : 0;
            g.sio[j] = si * row_bytes;
            g.rx[j] = p.x - qx; g.ry[j] = p.y - qy; g.rz[j] = p.z - qz;
        }
        g.any[j] = __ballot_sync(0xffffffffu, si < Ns);
    }
    return __reduce_add_sync(0xffffffffu, nn);
}

// influence of kernel point kpk on this lane's neighbour j: w = max(0, 1 - d/extent); in = inside the extent
template <int NH>
__device__ __forceinline__ float influence(const RowGeom<NH>& g, int j, const float4 kpk, float ext2, float inv_ext, bool& in) {
    const float ddx = g.rx[j] - kpk.x, ddy = g.ry[j] - kpk.y, ddz = g.rz[j] - kpk.z;
    const float d2 = ddx * ddx + ddy * ddy + ddz * ddz;
    in = d2 < ext2;
    return fmaxf(1.0f - sqrt_approx(d2) * inv_ext, 0.f);
}


// CSR influence list of ONE query row, built by a whole warp (lanes = neighbours), kernel-point-major so that the
// (byte offset of the feature row, weight) pairs of kernel point k follow those of k-1: ent[0..min(total, ECAP)) and
// off[0..KP_MAX_K] (off[K..] = total). Entries beyond ECAP are dropped; the caller checks off[KP_MAX_K] > ECAP.
// W16: the weight is stored as fp16 bits (round to nearest even) in the low half of .y — the form FHFMA consumes.
template <int NH, int ECAP, bool W16 = false>
__device__ __forceinline__ void build_row_list(const RowGeom<NH>& g, const float4* s_kp, int K, float ext2, float inv_ext,
                                               int2* ent, int* off, int lane) {
    const unsigned ltmask = (1u << lane) - 1u;
    int run = 0, myoff = 0;
#pragma unroll
    for (int k = 0; k < KP_MAX_K; ++k) {
        if (lane == k) myoff = run;
        if (k < K) {
            const float4 kpk = s_kp[k];
#pragma unroll
            for (int j = 0; j < NH; ++j) {
                if (g.any[j]) {                              // neighbours are distance-sorted: the tail group is often all pad
                    bool in;
                    const float w = influence<NH>(g, j, kpk, ext2, inv_ext, in);
                    const unsigned m = __ballot_sync(0xffffffffu, in);
                    const int pos = run + __popc(m & ltmask);
                    if (in && pos < ECAP) ent[pos] = make_int2(g.sio[j], W16 ? (int)__half_as_ushort(__float2half_rn(w)) : __float_as_int(w));
                    run += __popc(m);
                }
            }
        }
    }
    if (lane >= K) myoff = run;                    // off[K..K_MAX] = total (unused kernel points have empty lists)
    if (lane <= KP_MAX_K) off[lane] = myoff;
}

}  // namespace aprb
