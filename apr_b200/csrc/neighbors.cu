// neighbors.cu — K2 (cell-list build) + K3 (warp-cooperative radius query with distance-sorted top-W selection).
//
// Replaces batch_nanoflann_neighbors() (/root/reference/Predator_APR/cpp_wrappers/cpp_neighbors/neighbors/
// neighbors.cpp:211-332) and the nanoflann semantics it relies on (cpp_utils/nanoflann/nanoflann.hpp): membership
// d2 < r2 strict (:249-253) with d2 = (dx*dx + dy*dy) + dz*dz in fp32 without FMA, dx = query - support (:432-440),
// r2 = radius*radius in fp32 (neighbors.cpp:226), rows ascending in d2 (:1286-1287), pad = total support count
// (neighbors.cpp:322-324). Ties in d2 are broken by ascending support index — the order of the reference's own
// batch_ordered_neighbors (neighbors.cpp:176-181); nanoflann leaves ties to an unstable std::sort.
//
// Not a KD-tree port: supports are counting-sorted into a uniform grid per cloud (cell edge >= r, so a 3x3x3 window
// covers the ball); one warp per query scans 9 x-contiguous cell rows, compacts in-radius candidates as 64-bit keys
// (d2 bits << 32 | global support index) into a shared-memory buffer, and bitonic-sorts/truncates it to the W best.
#include "common.cuh"

namespace aprb {

struct NbGrid {
    float ox, oy, oz, inv_cell;  // origin corner (= support bbox min) and 1/cell edge
    float cell;
    int dx, dy, dz;              // cells per axis
    int base;                    // offset of this cloud's cells in the global cell arrays
};

// Cells budget per cloud: 16 cells per support point + 64 (LiDAR clouds are surfaces: most cells are empty).
__host__ __device__ __forceinline__ long long cell_share(int len) { return 16LL * len + 64; }

// single block, thread-strided over clouds; then a sequential prefix by thread 0 (B is small)
__global__ void nb_grid_params_kernel(const int* __restrict__ bbox, const int* __restrict__ soff, int B, float radius,
                                      NbGrid* __restrict__ grids, int* __restrict__ total_cells) {
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        NbGrid g;
        g.ox = g.oy = g.oz = 0.f; g.cell = radius; g.inv_cell = 0.f; g.dx = g.dy = g.dz = 0; g.base = 0;
        int len = soff[b + 1] - soff[b];
        if (len > 0) {
            const int* bb = bbox + 6 * b;
            g.ox = ord2f(bb[0]); g.oy = ord2f(bb[1]); g.oz = ord2f(bb[2]);
            float ex = ord2f(bb[3]) - g.ox, ey = ord2f(bb[4]) - g.oy, ez = ord2f(bb[5]) - g.oz;
            // cell edge slightly above r: the margin (2^-9 of a cell) dominates fp32 rounding of (p - o) / cell for
            // up to 2048 cells per axis, so a support within r of a query is never two cells away.
            float cell = radius * (1.0f + 1.0f / 512.0f);
            if (!(cell > 0.f)) cell = 1e-30f;
            long long share = cell_share(len);
            for (int it = 0; it < 200; ++it) {
                double nx = floor((double)ex / cell) + 1.0, ny = floor((double)ey / cell) + 1.0,
                       nz = floor((double)ez / cell) + 1.0;
                if (nx <= 2048.0 && ny <= 2048.0 && nz <= 2048.0 && nx * ny * nz <= (double)share) {
                    g.dx = (int)nx; g.dy = (int)ny; g.dz = (int)nz;
                    break;
                }
                cell *= 1.25992105f;  // coarsen: halves the cell count
            }
            if (g.dx == 0) { g.dx = g.dy = g.dz = 1; cell = 3.0e38f; }
            g.cell = cell;
            g.inv_cell = 1.0f / cell;
        }
        grids[b] = g;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < B; ++b) {
            grids[b].base = acc;
            acc += grids[b].dx * grids[b].dy * grids[b].dz;
        }
        *total_cells = acc;
    }
}

__device__ __forceinline__ int cell_coord(float p, float o, float inv_cell) {
    // unclamped cell coordinate; saturate to a safe int range so far-away queries cannot overflow
    float c = floorf((p - o) * inv_cell);
    c = fminf(fmaxf(c, -4.0f), 4100.0f);
    return (int)c;
}

__global__ void nb_count_kernel(const float* __restrict__ s, int Ns, const int* __restrict__ soff, int B,
                                const NbGrid* __restrict__ grids, int* __restrict__ cell_of,
                                int* __restrict__ cell_count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ns) return;
    int b = find_cloud(soff, B, i);
    NbGrid g = grids[b];
    int cx = min(max(cell_coord(s[3 * (size_t)i], g.ox, g.inv_cell), 0), g.dx - 1);
    int cy = min(max(cell_coord(s[3 * (size_t)i + 1], g.oy, g.inv_cell), 0), g.dy - 1);
    int cz = min(max(cell_coord(s[3 * (size_t)i + 2], g.oz, g.inv_cell), 0), g.dz - 1);
    int cell = g.base + (cz * g.dy + cy) * g.dx + cx;
    cell_of[i] = cell;
    atomicAdd(cell_count + cell, 1);
}

// sorted[pos] = (x, y, z, bits(global support index)); slot order inside a cell is arbitrary (the final
// (d2, index) sort makes the result deterministic)
__global__ void nb_scatter_kernel(const float* __restrict__ s, int Ns, const int* __restrict__ cell_of,
                                  const int* __restrict__ cell_start, int* __restrict__ cell_count,
                                  float4* __restrict__ sorted) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ns) return;
    int cell = cell_of[i];
    int pos = cell_start[cell] + atomicSub(cell_count + cell, 1) - 1;
    sorted[pos] = make_float4(s[3 * (size_t)i], s[3 * (size_t)i + 1], s[3 * (size_t)i + 2], __int_as_float(i));
}

__device__ __forceinline__ void warp_bitonic_sort(uint64_t* buf, int n, int lane) {
    // n is a power of two >= 32; ascending
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (n >> 1); t += 32) {
                int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                int p = i | j;
                uint64_t a = buf[i], b = buf[p];
                bool up = (i & k) == 0;
                if ((a > b) == up) { buf[i] = b; buf[p] = a; }
            }
            __syncwarp();
        }
    }
}

// The same network on registers: element i = slot * 32 + lane, E = n / 32 slots per lane. Exchanges at distance < 32
// are warp shuffles, larger ones swap between a lane's own slots. No shared-memory traffic, no bank conflicts (the
// shared-memory version was half of nb_query's instructions: 25 per compare-exchange step, ncu round 1).
template <int E>
__device__ __forceinline__ void warp_bitonic_sort_regs(uint64_t (&key)[E], int lane) {
    constexpr int n = 32 * E;
#pragma unroll
    for (int k = 2; k <= n; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int js = j >> 5;
#pragma unroll
                for (int s = 0; s < E; ++s) {
                    if ((s & js) == 0) {
                        const int p = s | js;
                        const bool up = (((s << 5) | lane) & k) == 0;
                        const uint64_t a = key[s], b = key[p];
                        if ((a > b) == up) { key[s] = b; key[p] = a; }
                    }
                }
            } else {
#pragma unroll
                for (int s = 0; s < E; ++s) {
                    const uint64_t other = __shfl_xor_sync(0xffffffffu, key[s], j);
                    const bool up = (((s << 5) | lane) & k) == 0;
                    const bool take_min = ((lane & j) == 0) == up;
                    if ((other < key[s]) == take_min) key[s] = other;
                }
            }
        }
    }
}

// final (d2, index) ordering of the `fill` keys in buf and the write of one output row
template <int E>
__device__ __forceinline__ void sort_and_emit(const uint64_t* buf, int fill, int W, int Ns, int* __restrict__ row, int lane) {
    uint64_t key[E];
#pragma unroll
    for (int s = 0; s < E; ++s) {
        const int t = s * 32 + lane;
        key[s] = t < fill ? buf[t] : ~0ull;
    }
    warp_bitonic_sort_regs<E>(key, lane);
    const int kept = min(fill, W);
#pragma unroll
    for (int s = 0; s < E; ++s) {
        const int t = s * 32 + lane;
        if (t < W) row[t] = t < kept ? (int)(uint32_t)key[s] : Ns;
    }
    for (int t = E * 32 + lane; t < W; t += 32) row[t] = Ns;
}

__device__ __forceinline__ int next_pow2_ge32(int v) {
    int n = 32;
    while (n < v) n <<= 1;
    return n;
}

// One warp per query. cap (power of two, >= W + 32) keys of shared memory per warp.
__global__ void __launch_bounds__(128)
nb_query_kernel(const float* __restrict__ q, int Nq, const int* __restrict__ qoff, const int* __restrict__ soff, int B,
                int Ns, const NbGrid* __restrict__ grids, const int* __restrict__ cell_start,
                const float4* __restrict__ sorted, float radius, int W, int cap, int* __restrict__ out, int ld,
                int* __restrict__ counts, int* __restrict__ max_count, int cps, int* __restrict__ seg_width) {
    extern __shared__ uint64_t s_keys[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int qi = blockIdx.x * (blockDim.x >> 5) + wib;
    if (qi >= Nq) return;
    uint64_t* buf = s_keys + (size_t)wib * cap;
    const float r2 = __fmul_rn(radius, radius);
    const int b = find_cloud(qoff, B, qi);
    const NbGrid g = grids[b];
    const float qx = q[3 * (size_t)qi], qy = q[3 * (size_t)qi + 1], qz = q[3 * (size_t)qi + 2];

    int fill = 0;            // keys currently in buf (warp-uniform)
    int total = 0;           // in-radius supports seen (warp-uniform)
    uint64_t thresh = ~0ull; // once truncated: the W-th best key so far

    if (g.dx > 0) {
        const int cx = cell_coord(qx, g.ox, g.inv_cell), cy = cell_coord(qy, g.oy, g.inv_cell),
                  cz = cell_coord(qz, g.oz, g.inv_cell);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dx - 1);
        int rs = 0, re = 0;  // lanes 0..8: candidate range of cell row (cy + r%3 - 1, cz + r/3 - 1)
        if (lane < 9 && x0 <= x1) {
            int yy = cy + (lane % 3) - 1, zz = cz + (lane / 3) - 1;
            if (yy >= 0 && yy < g.dy && zz >= 0 && zz < g.dz) {
                int row = g.base + (zz * g.dy + yy) * g.dx;
                rs = cell_start[row + x0];
                re = cell_start[row + x1 + 1];
            }
        }
        for (int r = 0; r < 9; ++r) {
            const int s0 = __shfl_sync(0xffffffffu, rs, r), e0 = __shfl_sync(0xffffffffu, re, r);
            for (int basej = s0; basej < e0; basej += 32) {
                const int j = basej + lane;
                bool hit = false;
                uint64_t key = 0;
                if (j < e0) {
                    const float4 sp = sorted[j];
                    const float ddx = __fsub_rn(qx, sp.x), ddy = __fsub_rn(qy, sp.y), ddz = __fsub_rn(qz, sp.z);
                    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)), __fmul_rn(ddz, ddz));
                    if (d2 < r2) {
                        hit = true;
                        key = ((uint64_t)__float_as_uint(d2) << 32) | (uint32_t)__float_as_int(sp.w);
                    }
                }
                const unsigned inr = __ballot_sync(0xffffffffu, hit);
                total += __popc(inr);
                bool keep = hit && key < thresh;
                unsigned m = __ballot_sync(0xffffffffu, keep);
                if (m) {
                    if (fill + 32 > cap) {  // make room: sort, keep the W best so far, tighten the threshold
                        for (int t = fill + lane; t < cap; t += 32) buf[t] = ~0ull;
                        __syncwarp();
                        warp_bitonic_sort(buf, cap, lane);
                        fill = min(fill, W);
                        thresh = buf[W - 1];  // sentinel (no filtering) while fewer than W keys are held
                        __syncwarp();
                        keep = keep && key < thresh;
                        m = __ballot_sync(0xffffffffu, keep);
                    }
                    if (keep) buf[fill + __popc(m & ((1u << lane) - 1))] = key;
                    fill += __popc(m);
                    __syncwarp();
                }
            }
        }
    }
    // final sort of what is left: in registers up to 128 keys, else in shared memory
    const int n2 = next_pow2_ge32(fill);
    int* row = out + (size_t)qi * ld;
    __syncwarp();
    if (n2 == 32) sort_and_emit<1>(buf, fill, W, Ns, row, lane);
    else if (n2 == 64) sort_and_emit<2>(buf, fill, W, Ns, row, lane);
    else if (n2 == 128) sort_and_emit<4>(buf, fill, W, Ns, row, lane);
    else {
        for (int t = fill + lane; t < n2; t += 32) buf[t] = ~0ull;
        __syncwarp();
        warp_bitonic_sort(buf, n2, lane);
        const int kept = min(fill, W);
        for (int t = lane; t < W; t += 32) row[t] = t < kept ? (int)(uint32_t)buf[t] : Ns;
    }
    if (lane == 0) {
        if (counts) counts[qi] = total;
        if (max_count) atomicMax(max_count, total);
        // width of the reference's matrix for this query's segment (collated pair): min(max_count, limit)
        if (seg_width && total > 0) atomicMax(seg_width + b / cps, min(total, W));
    }
}

// Nearest support only — column 0 of the matrix nb_query_kernel would produce (same fp32 non-FMA d2, strict d2 < r2, ties by
// ascending support index), pad = Ns when the ball is empty. This is all the reference ever reads of an upsample matrix
// (closest_pool: inds[:, 0], models/blocks.py:71-83). One thread per query: no candidate buffer, no sort.
__global__ void __launch_bounds__(128)
nb_nearest_kernel(const float* __restrict__ q, int Nq, const int* __restrict__ qoff, int B, int Ns,
                  const NbGrid* __restrict__ grids, const int* __restrict__ cell_start, const float4* __restrict__ sorted,
                  float radius, int* __restrict__ out, int ld) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= Nq) return;
    const float r2 = __fmul_rn(radius, radius);
    const int b = find_cloud(qoff, B, qi);
    const NbGrid g = grids[b];
    const float qx = q[3 * (size_t)qi], qy = q[3 * (size_t)qi + 1], qz = q[3 * (size_t)qi + 2];
    uint64_t best = ~0ull;
    if (g.dx > 0) {
        const int cx = cell_coord(qx, g.ox, g.inv_cell), cy = cell_coord(qy, g.oy, g.inv_cell),
                  cz = cell_coord(qz, g.oz, g.inv_cell);
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dx - 1);
        if (x0 <= x1) {
            for (int zz = max(cz - 1, 0); zz <= min(cz + 1, g.dz - 1); ++zz) {
                for (int yy = max(cy - 1, 0); yy <= min(cy + 1, g.dy - 1); ++yy) {
                    const int row = g.base + (zz * g.dy + yy) * g.dx;
                    const int s0 = cell_start[row + x0], e0 = cell_start[row + x1 + 1];
                    for (int j = s0; j < e0; ++j) {
                        const float4 sp = sorted[j];
                        const float ddx = __fsub_rn(qx, sp.x), ddy = __fsub_rn(qy, sp.y), ddz = __fsub_rn(qz, sp.z);
                        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)), __fmul_rn(ddz, ddz));
                        if (d2 < r2) {
                            const uint64_t key = ((uint64_t)__float_as_uint(d2) << 32) | (uint32_t)__float_as_int(sp.w);
                            best = key < best ? key : best;
                        }
                    }
                }
            }
        }
    }
    out[(size_t)qi * ld] = best == ~0ull ? Ns : (int)(uint32_t)best;
}

struct NbWs {
    int *qoff, *soff, *bbox, *total_cells, *cell_of, *cell_count, *cell_start;
    NbGrid* grids;
    float4* sorted;
    void* temp;
    size_t temp_bytes;
    long long cells_cap;
};

static size_t carve_nb(Carver& c, int Nq, int Ns, int B, NbWs* w) {
    (void)Nq;
    NbWs t;
    t.cells_cap = 16LL * Ns + 64LL * B + 1;
    t.qoff = c.take<int>(B + 1);
    t.soff = c.take<int>(B + 1);
    t.bbox = c.take<int>(6 * (size_t)B);
    t.total_cells = c.take<int>(4);
    t.grids = c.take<NbGrid>(B);
    t.cell_of = c.take<int>(Ns);
    t.cell_count = c.take<int>((size_t)t.cells_cap + 1);
    t.cell_start = c.take<int>((size_t)t.cells_cap + 1);
    t.sorted = c.take<float4>(Ns);
    t.temp_bytes = scan_temp_bytes((int)t.cells_cap + 1);
    t.temp = c.take<char>(t.temp_bytes);
    if (w) *w = t;
    return c.off;
}

}  // namespace aprb

using namespace aprb;

static int build_grid(const float* d_s, const int32_t* d_slens, int B, int Ns, float radius, NbWs& w, cudaStream_t st) {
    const int T = 256;
    APRB_TIMED("setup_kernel", st, 1, (setup_kernel<<<1, 256, 0, st>>>(d_slens, w.soff, nullptr, nullptr, B, w.bbox, nullptr, 0)));
    if (Ns > 0) APRB_TIMED("bbox_kernel", st, 1, (bbox_kernel<<<cdiv(Ns, T), T, 0, st>>>(d_s, Ns, w.soff, B, w.bbox)));
    APRB_TIMED("nb_grid_params_kernel", st, 1, (nb_grid_params_kernel<<<1, 256, 0, st>>>(w.bbox, w.soff, B, radius, w.grids, w.total_cells)));
    APRB_CUDA_OK(cudaMemsetAsync(w.cell_count, 0, sizeof(int) * ((size_t)w.cells_cap + 1), st));
    if (Ns > 0) APRB_TIMED("nb_count_kernel", st, 1, (nb_count_kernel<<<cdiv(Ns, T), T, 0, st>>>(d_s, Ns, w.soff, B, w.grids, w.cell_of, w.cell_count)));
    APRB_LAUNCH_OK();
    int rc = exclusive_scan_i32(w.cell_count, w.cell_start, (int)w.cells_cap + 1, w.temp, w.temp_bytes, st);
    if (rc) return rc;
    if (Ns > 0) APRB_TIMED("nb_scatter_kernel", st, 1, (nb_scatter_kernel<<<cdiv(Ns, T), T, 0, st>>>(d_s, Ns, w.cell_of, w.cell_start, w.cell_count, w.sorted)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

static int query_grid(const float* d_q, const int32_t* d_qlens, int B, int Nq, int Ns, float radius, int width,
                      int32_t* d_out_idx, int ld, int32_t* d_counts, int32_t* d_max_count, NbWs& w, cudaStream_t st,
                      int cps = 1, int32_t* d_seg_width = nullptr) {
    if (d_seg_width) APRB_CUDA_OK(cudaMemsetAsync(d_seg_width, 0, sizeof(int) * (size_t)cdiv(B, cps), st));
    // query offsets + reset of the max_count scalar in one tiny launch
    APRB_TIMED("setup_kernel", st, 1, (setup_kernel<<<1, 256, 0, st>>>(d_qlens, w.qoff, nullptr, nullptr, B, nullptr, d_max_count, d_max_count ? 1 : 0)));
    int cap = 64;
    while (cap < width + 32) cap <<= 1;
    int wpb = 4;  // warps per block
    while (wpb > 1 && (size_t)wpb * cap * sizeof(uint64_t) > 200 * 1024) wpb >>= 1;
    size_t smem = (size_t)wpb * cap * sizeof(uint64_t);
    if (smem > 48 * 1024)
        APRB_CUDA_OK(cudaFuncSetAttribute(nb_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    APRB_TIMED("nb_query_kernel", st, 1, (nb_query_kernel<<<cdiv(Nq, wpb), wpb * 32, smem, st>>>(d_q, Nq, w.qoff, w.soff, B, Ns, w.grids, w.cell_start, w.sorted,
                                                        radius, width, cap, d_out_idx, ld, d_counts, d_max_count, cps, d_seg_width)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

extern "C" size_t aprb_radius_neighbors_ws_bytes(int Nq, int Ns, int B) {
    if (Nq < 0 || Ns < 0 || B < 0) return 0;
    Carver c(nullptr, 0);
    return carve_nb(c, Nq, Ns > 0 ? Ns : 1, B > 0 ? B : 1, nullptr) + 256;
}

extern "C" size_t aprb_cell_grid_bytes(int Ns, int B) { return aprb_radius_neighbors_ws_bytes(0, Ns, B); }

#define NB_COMMON_CHECKS()                                                                                            \
    APRB_REQUIRE(B >= 1 && Ns >= 0, "need B >= 1, Ns >= 0");                                                          \
    APRB_REQUIRE((long long)Ns * 16 + 64LL * B < 0x7FFFFF00LL, "support set too large for the int32 cell table")

extern "C" int aprb_cell_grid_build(const float* d_s, const int32_t* d_slens, int B, int Ns, float radius, void* d_grid,
                                    size_t grid_bytes, void* stream) {
    NB_COMMON_CHECKS();
    APRB_REQUIRE(d_slens && d_grid && (d_s || Ns == 0), "null pointer");
    APRB_REQUIRE(radius > 0.f, "radius must be positive");
    Carver c(d_grid, grid_bytes);
    NbWs w;
    carve_nb(c, 0, Ns > 0 ? Ns : 1, B, &w);
    if (!c.ok()) { set_error("aprb_cell_grid_build: grid buffer too small (%zu < %zu)", grid_bytes, c.off); return APRB_ERR_WORKSPACE; }
    return build_grid(d_s, d_slens, B, Ns, radius, w, (cudaStream_t)stream);
}

extern "C" int aprb_cell_grid_query(const void* d_grid, size_t grid_bytes, const float* d_q, const int32_t* d_qlens, int B,
                                    int Nq, int Ns, float radius, int width, int32_t* d_out_idx, int ld,
                                    int32_t* d_counts, int32_t* d_max_count, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NB_COMMON_CHECKS();
    APRB_REQUIRE(Nq >= 0 && d_qlens && d_grid, "bad query arguments");
    APRB_REQUIRE(width >= 1 && width <= 16352, "width must be in [1, 16352]");
    APRB_REQUIRE(ld >= width, "ld < width");
    if (Nq == 0) { if (d_max_count) APRB_CUDA_OK(cudaMemsetAsync(d_max_count, 0, sizeof(int), st)); return APRB_OK; }
    APRB_REQUIRE(d_q && d_out_idx, "null point/output pointer");
    Carver c(const_cast<void*>(d_grid), grid_bytes);
    NbWs w;
    carve_nb(c, 0, Ns > 0 ? Ns : 1, B, &w);
    if (!c.ok()) { set_error("aprb_cell_grid_query: grid buffer too small"); return APRB_ERR_WORKSPACE; }
    return query_grid(d_q, d_qlens, B, Nq, Ns, radius, width, d_out_idx, ld, d_counts, d_max_count, w, st);
}

// aprb_cell_grid_query that also records, per segment of `clouds_per_segment` clouds, the width the reference's matrix
// would have for that collate: d_seg_width[s] = min(max over the segment's queries of the neighbour count, width).
extern "C" int aprb_cell_grid_query_seg(const void* d_grid, size_t grid_bytes, const float* d_q, const int32_t* d_qlens, int B,
                                        int Nq, int Ns, float radius, int width, int32_t* d_out_idx, int ld,
                                        int clouds_per_segment, int32_t* d_seg_width, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NB_COMMON_CHECKS();
    APRB_REQUIRE(Nq >= 0 && d_qlens && d_grid && d_seg_width && clouds_per_segment >= 1, "bad query arguments");
    APRB_REQUIRE(width >= 1 && width <= 16352, "width must be in [1, 16352]");
    APRB_REQUIRE(ld >= width, "ld < width");
    if (Nq == 0) { APRB_CUDA_OK(cudaMemsetAsync(d_seg_width, 0, sizeof(int) * (size_t)cdiv(B, clouds_per_segment), st)); return APRB_OK; }
    APRB_REQUIRE(d_q && d_out_idx, "null point/output pointer");
    Carver c(const_cast<void*>(d_grid), grid_bytes);
    NbWs w;
    carve_nb(c, 0, Ns > 0 ? Ns : 1, B, &w);
    if (!c.ok()) { set_error("aprb_cell_grid_query_seg: grid buffer too small"); return APRB_ERR_WORKSPACE; }
    return query_grid(d_q, d_qlens, B, Nq, Ns, radius, width, d_out_idx, ld, nullptr, nullptr, w, st, clouds_per_segment, d_seg_width);
}

// Nearest support within `radius` per query: out[n * ld] = column 0 of aprb_cell_grid_query's matrix (pad Ns).
extern "C" int aprb_cell_grid_query_nearest(const void* d_grid, size_t grid_bytes, const float* d_q, const int32_t* d_qlens, int B,
                                            int Nq, int Ns, float radius, int32_t* d_out_idx, int ld, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NB_COMMON_CHECKS();
    APRB_REQUIRE(Nq >= 0 && d_qlens && d_grid && ld >= 1, "bad query arguments");
    if (Nq == 0) return APRB_OK;
    APRB_REQUIRE(d_q && d_out_idx, "null point/output pointer");
    Carver c(const_cast<void*>(d_grid), grid_bytes);
    NbWs w;
    carve_nb(c, 0, Ns > 0 ? Ns : 1, B, &w);
    if (!c.ok()) { set_error("aprb_cell_grid_query_nearest: grid buffer too small"); return APRB_ERR_WORKSPACE; }
    APRB_TIMED("setup_kernel", st, 1, (setup_kernel<<<1, 256, 0, st>>>(d_qlens, w.qoff, nullptr, nullptr, B, nullptr, nullptr, 0)));
    APRB_TIMED("nb_nearest_kernel", st, 1, (nb_nearest_kernel<<<cdiv(Nq, 128), 128, 0, st>>>(d_q, Nq, w.qoff, B, Ns, w.grids, w.cell_start, w.sorted,
                                                                                             radius, d_out_idx, ld)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

extern "C" int aprb_radius_neighbors_batch(const float* d_q, const float* d_s, const int32_t* d_qlens,
                                           const int32_t* d_slens, int B, int Nq, int Ns, float radius, int width,
                                           int32_t* d_out_idx, int ld, int32_t* d_counts, int32_t* d_max_count,
                                           void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    NB_COMMON_CHECKS();
    APRB_REQUIRE(Nq >= 0 && d_qlens && d_slens, "null batch-length pointer");
    APRB_REQUIRE(width >= 1 && width <= 16352, "width must be in [1, 16352]");
    APRB_REQUIRE(ld >= width, "ld < width");
    if (Nq == 0) { if (d_max_count) APRB_CUDA_OK(cudaMemsetAsync(d_max_count, 0, sizeof(int), st)); return APRB_OK; }
    APRB_REQUIRE(d_q && d_out_idx && d_ws && (d_s || Ns == 0), "null point/output/workspace pointer");
    Carver c(d_ws, ws_bytes);
    NbWs w;
    carve_nb(c, Nq, Ns > 0 ? Ns : 1, B, &w);
    if (!c.ok()) { set_error("aprb_radius_neighbors_batch: workspace too small (%zu < %zu)", ws_bytes, c.off); return APRB_ERR_WORKSPACE; }
    int rc = build_grid(d_s, d_slens, B, Ns, radius, w, st);
    if (rc) return rc;
    return query_grid(d_q, d_qlens, B, Nq, Ns, radius, width, d_out_idx, ld, d_counts, d_max_count, w, st);
}
