// subsample.cu — K1: batched voxel-grid barycentre subsampling.
//
// Replaces grid_subsampling()/batch_grid_subsampling()
// (/root/reference/Predator_APR/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106, :109-211).
// Arithmetic contract reproduced bit-for-bit: origin = floor(min * (1/dl)) * dl in fp32 (:27), voxel index
// floor((p - origin) / dl) with a true fp32 division (:53-55), per-voxel fp32 sum in INPUT ORDER
// (grid_subsampling.h:74-79), barycentre = sum * float(1.0 / count) (:87). The reference's row order is
// std::unordered_map iteration order (unspecified); we emit ascending (cloud, iz, iy, ix), which is the same total
// order as ascending (cloud, reference key) because key = ix + NX*iy + NX*NY*iz with ix < NX, iy < NY.
//
// Pipeline (all stream-ordered, no host sync): offsets -> per-cloud bbox (ordered-int atomics) -> per-cloud grid
// params -> 64-bit packed keys -> stable radix sort (key, point index) -> head flags -> exclusive scan -> per-cloud
// lengths (+max_p) -> one thread per voxel sums its members sequentially in ascending point index and emits.
#include "common.cuh"

namespace aprb {

struct CloudGrid {
    float ox, oy, oz;  // origin corner
    int nx, ny, nz;    // grid dims (voxels per axis)
};

// one thread per cloud; also reduces the global max dims (for the packed-key bit widths)
__global__ void sub_params_kernel(const int* __restrict__ bbox, const int* __restrict__ off, int B, float dl,
                                  CloudGrid* __restrict__ grids, int* __restrict__ gdims) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    CloudGrid g = {0.f, 0.f, 0.f, 0, 0, 0};
    if (off[b + 1] > off[b]) {
        const int* bb = bbox + 6 * b;
        float inv = __fdiv_rn(1.0f, dl);
        g.ox = __fmul_rn(floorf(__fmul_rn(ord2f(bb[0]), inv)), dl);
        g.oy = __fmul_rn(floorf(__fmul_rn(ord2f(bb[1]), inv)), dl);
        g.oz = __fmul_rn(floorf(__fmul_rn(ord2f(bb[2]), inv)), dl);
        long long nx = (long long)floorf(__fdiv_rn(__fsub_rn(ord2f(bb[3]), g.ox), dl)) + 1;
        long long ny = (long long)floorf(__fdiv_rn(__fsub_rn(ord2f(bb[4]), g.oy), dl)) + 1;
        long long nz = (long long)floorf(__fdiv_rn(__fsub_rn(ord2f(bb[5]), g.oz), dl)) + 1;
        const long long lim = 0x7FFFFFF0LL;
        g.nx = (int)min(max(nx, 1LL), lim); g.ny = (int)min(max(ny, 1LL), lim); g.nz = (int)min(max(nz, 1LL), lim);
        atomicMax(gdims + 0, g.nx); atomicMax(gdims + 1, g.ny); atomicMax(gdims + 2, g.nz);
    }
    grids[b] = g;
}

__device__ __forceinline__ int bitlen(unsigned v) { return 32 - __clz(v); }  // bits to represent v (0 -> 0)

template <typename KeyT>
__global__ void sub_keys_kernel(const float* __restrict__ pts, int N, const int* __restrict__ off, int B, float dl,
                                const CloudGrid* __restrict__ grids, const int* __restrict__ gdims,
                                KeyT* __restrict__ keys, int* __restrict__ vals, int* __restrict__ status) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int bx = bitlen((unsigned)max(gdims[0] - 1, 0)), by = bitlen((unsigned)max(gdims[1] - 1, 0)),
        bz = bitlen((unsigned)max(gdims[2] - 1, 0)), bc = bitlen((unsigned)max(B - 1, 0));
    // status: 0 = ok, 1 = grid does not fit 64 bits at all, 2 = needs the 64-bit key path (caller re-runs with key_bits=64)
    if (i == 0 && status) *status = (bx + by + bz + bc > 64) ? 1 : ((bx + by + bz + bc > (int)sizeof(KeyT) * 8) ? 2 : 0);
    int b = find_cloud(off, B, i);
    CloudGrid g = grids[b];
    float px = pts[3 * (size_t)i], py = pts[3 * (size_t)i + 1], pz = pts[3 * (size_t)i + 2];
    long long ix = (long long)floorf(__fdiv_rn(__fsub_rn(px, g.ox), dl));
    long long iy = (long long)floorf(__fdiv_rn(__fsub_rn(py, g.oy), dl));
    long long iz = (long long)floorf(__fdiv_rn(__fsub_rn(pz, g.oz), dl));
    // p >= min >= origin up to one rounding; clamp the (measure-zero) out-of-range case instead of wrapping
    ix = min(max(ix, 0LL), (long long)g.nx - 1); iy = min(max(iy, 0LL), (long long)g.ny - 1);
    iz = min(max(iz, 0LL), (long long)g.nz - 1);
    uint64_t key = (uint64_t)b;
    key = (key << bz) | (uint64_t)iz;
    key = (key << by) | (uint64_t)iy;
    key = (key << bx) | (uint64_t)ix;
    keys[i] = (KeyT)key;
    vals[i] = i;
}

// flags[i] = 1 iff sorted position i starts a new voxel; flags[N] = 0 (so the exclusive scan's entry N is the total)
template <typename KeyT>
__global__ void sub_flags_kernel(const KeyT* __restrict__ skeys, int N, int* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > N) return;
    flags[i] = (i < N) && (i == 0 || skeys[i] != skeys[i - 1]);
}

// single block: per-cloud voxel counts -> output lengths (max_p) -> output row starts; total M.
// The sorted array keeps clouds contiguous with the input's offsets (cloud id is the most significant key field).
__global__ void sub_lens_kernel(const int* __restrict__ pos, const int* __restrict__ off, int B, int max_p,
                                int* __restrict__ out_lens, int* __restrict__ out_start, int* __restrict__ out_M) {
    __shared__ int s_part[256];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < B; base += blockDim.x) {
        int b = base + threadIdx.x;
        int v = 0;
        if (b < B) {
            v = pos[off[b + 1]] - pos[off[b]];
            if (max_p > 0 && v > max_p) v = max_p;
            out_lens[b] = v;
        }
        s_part[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < blockDim.x; d <<= 1) {
            int t = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0;
            __syncthreads();
            s_part[threadIdx.x] += t;
            __syncthreads();
        }
        int incl = s_part[threadIdx.x], carry = s_carry;
        if (b < B) out_start[b] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *out_M = s_carry;
}

template <typename KeyT>
__global__ void sub_emit_kernel(const float* __restrict__ pts, const float* __restrict__ feats, int fdim, int N,
                                const KeyT* __restrict__ skeys, const int* __restrict__ svals,
                                const int* __restrict__ pos, const int* __restrict__ off, int B, int max_p,
                                const int* __restrict__ out_start, float* __restrict__ out_pts,
                                float* __restrict__ out_feats) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    KeyT key = skeys[i];
    if (i > 0 && skeys[i - 1] == key) return;  // not a voxel head
    int b = find_cloud(off, B, i);
    int rank = pos[i] - pos[off[b]];            // voxel rank inside its cloud (canonical order)
    if (max_p > 0 && rank >= max_p) return;
    int row = out_start[b] + rank;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    int cnt = 0;
    int end = off[b + 1];
    for (int j = i; j < end && skeys[j] == key; ++j) {   // stable sort => ascending original index = input order
        int p = svals[j];
        sx = __fadd_rn(sx, pts[3 * (size_t)p]); sy = __fadd_rn(sy, pts[3 * (size_t)p + 1]);
        sz = __fadd_rn(sz, pts[3 * (size_t)p + 2]);
        ++cnt;
    }
    float w = (float)(1.0 / (double)cnt);
    out_pts[3 * (size_t)row] = __fmul_rn(sx, w); out_pts[3 * (size_t)row + 1] = __fmul_rn(sy, w);
    out_pts[3 * (size_t)row + 2] = __fmul_rn(sz, w);
    if (feats) {   // per-voxel feature mean: sum in input order, then f / count (grid_subsampling.cpp:90-95)
        float c = (float)cnt;
        for (int d = 0; d < fdim; ++d) {
            float s = 0.f;
            for (int j = i; j < i + cnt; ++j) s = __fadd_rn(s, feats[(size_t)svals[j] * fdim + d]);
            out_feats[(size_t)row * fdim + d] = __fdiv_rn(s, c);
        }
    }
}

// Per-voxel label vote (grid_subsampling.cpp:63-68 update_classes, :96-101 max_element over the per-dimension
// label -> count map): out label = the most frequent label of the voxel's points, per label dimension. The reference
// breaks ties by unordered_map iteration order (unspecified); here the smallest label value wins. One thread per
// voxel head, O(n^2) over the voxel's points (a voxel holds a handful of points; off the hot path).
template <typename KeyT>
__global__ void sub_vote_kernel(const int* __restrict__ classes, int ldim, int N, const KeyT* __restrict__ skeys,
                                const int* __restrict__ svals, const int* __restrict__ pos, const int* __restrict__ off,
                                int B, int max_p, const int* __restrict__ out_start, int* __restrict__ out_classes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    KeyT key = skeys[i];
    if (i > 0 && skeys[i - 1] == key) return;
    int b = find_cloud(off, B, i);
    int rank = pos[i] - pos[off[b]];
    if (max_p > 0 && rank >= max_p) return;
    int row = out_start[b] + rank;
    int end = off[b + 1], cnt = 0;
    for (int j = i; j < end && skeys[j] == key; ++j) ++cnt;
    for (int d = 0; d < ldim; ++d) {
        int best = 0, bestc = 0;
        for (int j = i; j < i + cnt; ++j) {
            const int lj = classes[(size_t)svals[j] * ldim + d];
            bool first = true;
            for (int jj = i; jj < j; ++jj)
                if (classes[(size_t)svals[jj] * ldim + d] == lj) { first = false; break; }
            if (!first) continue;
            int c = 1;
            for (int jj = j + 1; jj < i + cnt; ++jj) c += classes[(size_t)svals[jj] * ldim + d] == lj ? 1 : 0;
            if (c > bestc || (c == bestc && lj < best)) { best = lj; bestc = c; }
        }
        out_classes[(size_t)row * ldim + d] = best;
    }
}

// ---- first-level voxelisation of raw scans, open3d semantics ------------------------------------------------------------
// The step before the path: raw KITTI scans are float32 [n, 4] (x, y, z, reflectance; datasets/kitti.py:191-194) and are
// voxelised with open3d's PointCloud.voxel_down_sample(voxel_size) (kitti.py:468-471, :588-589; open3d==0.10.0.0,
// requirements.txt:5 — third party, absent here: parity UNPINNED, the algorithm below restates its published source,
// open3d/geometry/PointCloud.cpp VoxelDownSample): points widened to double, voxel_min_bound = min_bound - 0.5 * voxel_size,
// voxel index = floor((p - voxel_min_bound) / voxel_size) per axis, per-voxel double sum in point order, output =
// sum / count. The reference then narrows to fp32 (dataloader.py:125/:163). Same arithmetic here (double index and sum,
// input-order summation, one final narrowing); the row order is ascending (cloud, iz, iy, ix) instead of open3d's
// unordered_map iteration order.
struct O3dGrid {
    double ox, oy, oz;
    int nx, ny, nz;
};

__global__ void raw_bbox_kernel(const float* __restrict__ pts, int stride, int N, const int* __restrict__ off, int B,
                                int* __restrict__ bbox) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int b = find_cloud(off, B, i);
    const float* p = pts + (size_t)i * stride;
    int* bb = bbox + 6 * b;
    // warp-level pre-reduction when the whole warp lies in one cloud
    int v[6] = {f2ord(p[0]), f2ord(p[1]), f2ord(p[2]), f2ord(p[0]), f2ord(p[1]), f2ord(p[2])};
    const unsigned act = __activemask();
    const int b0 = __shfl_sync(act, b, __ffs(act) - 1);
    if (__all_sync(act, b == b0)) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { v[k] = __reduce_min_sync(act, v[k]); v[3 + k] = __reduce_max_sync(act, v[3 + k]); }
        if ((threadIdx.x & 31) == __ffs(act) - 1) {
            for (int k = 0; k < 3; ++k) { atomicMin(bb + k, v[k]); atomicMax(bb + 3 + k, v[3 + k]); }
        }
    } else {
        for (int k = 0; k < 3; ++k) { atomicMin(bb + k, v[k]); atomicMax(bb + 3 + k, v[3 + k]); }
    }
}

__global__ void o3d_params_kernel(const int* __restrict__ bbox, const int* __restrict__ off, int B, double vs,
                                  O3dGrid* __restrict__ grids, int* __restrict__ gdims) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    O3dGrid g = {0.0, 0.0, 0.0, 0, 0, 0};
    if (off[b + 1] > off[b]) {
        const int* bb = bbox + 6 * b;
        g.ox = (double)ord2f(bb[0]) - vs * 0.5; g.oy = (double)ord2f(bb[1]) - vs * 0.5; g.oz = (double)ord2f(bb[2]) - vs * 0.5;
        const double lim = 2147483632.0;
        g.nx = (int)fmin(fmax(floor(((double)ord2f(bb[3]) - g.ox) / vs) + 1.0, 1.0), lim);
        g.ny = (int)fmin(fmax(floor(((double)ord2f(bb[4]) - g.oy) / vs) + 1.0, 1.0), lim);
        g.nz = (int)fmin(fmax(floor(((double)ord2f(bb[5]) - g.oz) / vs) + 1.0, 1.0), lim);
        atomicMax(gdims + 0, g.nx); atomicMax(gdims + 1, g.ny); atomicMax(gdims + 2, g.nz);
    }
    grids[b] = g;
}

template <typename KeyT>
__global__ void o3d_keys_kernel(const float* __restrict__ pts, int stride, int N, const int* __restrict__ off, int B, double vs,
                                const O3dGrid* __restrict__ grids, const int* __restrict__ gdims, KeyT* __restrict__ keys,
                                int* __restrict__ vals, int* __restrict__ status) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int bx = bitlen((unsigned)max(gdims[0] - 1, 0)), by = bitlen((unsigned)max(gdims[1] - 1, 0)),
        bz = bitlen((unsigned)max(gdims[2] - 1, 0)), bc = bitlen((unsigned)max(B - 1, 0));
    if (i == 0 && status) *status = (bx + by + bz + bc > 64) ? 1 : ((bx + by + bz + bc > (int)sizeof(KeyT) * 8) ? 2 : 0);
    int b = find_cloud(off, B, i);
    O3dGrid g = grids[b];
    const float* p = pts + (size_t)i * stride;
    long long ix = (long long)floor(((double)p[0] - g.ox) / vs), iy = (long long)floor(((double)p[1] - g.oy) / vs),
              iz = (long long)floor(((double)p[2] - g.oz) / vs);
    ix = min(max(ix, 0LL), (long long)g.nx - 1); iy = min(max(iy, 0LL), (long long)g.ny - 1);
    iz = min(max(iz, 0LL), (long long)g.nz - 1);
    uint64_t key = (uint64_t)b;
    key = (key << bz) | (uint64_t)iz;
    key = (key << by) | (uint64_t)iy;
    key = (key << bx) | (uint64_t)ix;
    keys[i] = (KeyT)key;
    vals[i] = i;
}

template <typename KeyT>
__global__ void o3d_emit_kernel(const float* __restrict__ pts, int stride, int N, const KeyT* __restrict__ skeys,
                                const int* __restrict__ svals, const int* __restrict__ pos, const int* __restrict__ off, int B,
                                const int* __restrict__ out_start, float* __restrict__ out_pts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    KeyT key = skeys[i];
    if (i > 0 && skeys[i - 1] == key) return;  // not a voxel head
    int b = find_cloud(off, B, i);
    int row = out_start[b] + pos[i] - pos[off[b]];
    double sx = 0.0, sy = 0.0, sz = 0.0;
    int cnt = 0;
    int end = off[b + 1];
    for (int j = i; j < end && skeys[j] == key; ++j) {   // stable sort => ascending point index = open3d's accumulation order
        const float* p = pts + (size_t)svals[j] * stride;
        sx += (double)p[0]; sy += (double)p[1]; sz += (double)p[2];
        ++cnt;
    }
    const double c = (double)cnt;
    out_pts[3 * (size_t)row] = (float)(sx / c); out_pts[3 * (size_t)row + 1] = (float)(sy / c);
    out_pts[3 * (size_t)row + 2] = (float)(sz / c);
}

struct SubWs {
    int *off, *bbox, *gdims, *vals_in, *vals_out, *flags, *pos, *out_start;
    CloudGrid* grids;
    uint64_t *keys_in, *keys_out;
    void* temp;
    size_t temp_bytes;
};

static size_t carve_sub(Carver& c, int N, int B, SubWs* w) {
    SubWs t;
    t.off = c.take<int>(B + 1);
    t.bbox = c.take<int>(6 * (size_t)B);
    t.gdims = c.take<int>(4);
    t.grids = c.take<CloudGrid>(B);
    t.out_start = c.take<int>(B + 1);
    t.keys_in = c.take<uint64_t>(N);
    t.keys_out = c.take<uint64_t>(N);
    t.vals_in = c.take<int>(N);
    t.vals_out = c.take<int>(N);
    t.flags = c.take<int>((size_t)N + 1);
    t.pos = c.take<int>((size_t)N + 1);
    size_t a = sort_temp_bytes(N), b = scan_temp_bytes(N + 1);
    t.temp_bytes = a > b ? a : b;
    t.temp = c.take<char>(t.temp_bytes);
    if (w) *w = t;
    return c.off;
}

}  // namespace aprb

using namespace aprb;

extern "C" size_t aprb_grid_subsample_ws_bytes(int N, int B, int fdim) {
    (void)fdim;
    if (N < 0 || B < 0) return 0;
    Carver c(nullptr, 0);
    return carve_sub(c, N > 0 ? N : 1, B > 0 ? B : 1, nullptr) + 256;
}

template <typename KeyT>
static int run_subsample(const float* d_pts, const int32_t* d_lens, int B, int N, float dl, int max_p, const float* d_feats,
                         int fdim, float* d_out_pts, int32_t* d_out_lens, int32_t* d_out_M, float* d_out_feats,
                         int32_t* d_status, SubWs& w, cudaStream_t st, const int32_t* d_classes = nullptr, int ldim = 0,
                         int32_t* d_out_classes = nullptr) {
    const int T = 256;
    KeyT* keys_in = reinterpret_cast<KeyT*>(w.keys_in);
    KeyT* keys_out = reinterpret_cast<KeyT*>(w.keys_out);
    APRB_TIMED("setup_kernel", st, 1, (setup_kernel<<<1, 256, 0, st>>>(d_lens, w.off, nullptr, nullptr, B, w.bbox, w.gdims, 4)));
    APRB_TIMED("bbox_kernel", st, 1, (bbox_kernel<<<cdiv(N, T), T, 0, st>>>(d_pts, N, w.off, B, w.bbox)));
    APRB_TIMED("sub_params_kernel", st, 1, (sub_params_kernel<<<cdiv(B, T), T, 0, st>>>(w.bbox, w.off, B, dl, w.grids, w.gdims)));
    APRB_TIMED("sub_keys_kernel", st, 1, (sub_keys_kernel<KeyT><<<cdiv(N, T), T, 0, st>>>(d_pts, N, w.off, B, dl, w.grids, w.gdims, keys_in, w.vals_in, d_status)));
    APRB_LAUNCH_OK();
    int rc = sort_pairs_i32(keys_in, keys_out, w.vals_in, w.vals_out, N, w.temp, w.temp_bytes, st);
    if (rc) return rc;
    APRB_TIMED("sub_flags_kernel", st, 1, (sub_flags_kernel<KeyT><<<cdiv(N + 1, T), T, 0, st>>>(keys_out, N, w.flags)));
    rc = exclusive_scan_i32(w.flags, w.pos, N + 1, w.temp, w.temp_bytes, st);
    if (rc) return rc;
    APRB_TIMED("sub_lens_kernel", st, 1, (sub_lens_kernel<<<1, 256, 0, st>>>(w.pos, w.off, B, max_p, d_out_lens, w.out_start, d_out_M)));
    APRB_TIMED("sub_emit_kernel", st, 1, (sub_emit_kernel<KeyT><<<cdiv(N, T), T, 0, st>>>(d_pts, d_feats, fdim, N, keys_out, w.vals_out, w.pos, w.off, B, max_p,
                                                    w.out_start, d_out_pts, d_out_feats)));
    if (d_classes)
        APRB_TIMED("sub_vote_kernel", st, 1, (sub_vote_kernel<KeyT><<<cdiv(N, T), T, 0, st>>>(d_classes, ldim, N, keys_out, w.vals_out, w.pos, w.off, B, max_p,
                                                        w.out_start, d_out_classes)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

template <typename KeyT>
static int run_voxel_raw(const float* d_raw, int stride, const int32_t* d_lens, int B, int N, double vs, float* d_out_pts,
                         int32_t* d_out_lens, int32_t* d_out_M, int32_t* d_status, SubWs& w, O3dGrid* grids, cudaStream_t st) {
    const int T = 256;
    KeyT* keys_in = reinterpret_cast<KeyT*>(w.keys_in);
    KeyT* keys_out = reinterpret_cast<KeyT*>(w.keys_out);
    APRB_TIMED("setup_kernel", st, 1, (setup_kernel<<<1, 256, 0, st>>>(d_lens, w.off, nullptr, nullptr, B, w.bbox, w.gdims, 4)));
    APRB_TIMED("raw_bbox_kernel", st, 1, (raw_bbox_kernel<<<cdiv(N, T), T, 0, st>>>(d_raw, stride, N, w.off, B, w.bbox)));
    APRB_TIMED("o3d_params_kernel", st, 1, (o3d_params_kernel<<<cdiv(B, T), T, 0, st>>>(w.bbox, w.off, B, vs, grids, w.gdims)));
    APRB_TIMED("o3d_keys_kernel", st, 1, (o3d_keys_kernel<KeyT><<<cdiv(N, T), T, 0, st>>>(d_raw, stride, N, w.off, B, vs, grids, w.gdims, keys_in, w.vals_in, d_status)));
    APRB_LAUNCH_OK();
    int rc = sort_pairs_i32(keys_in, keys_out, w.vals_in, w.vals_out, N, w.temp, w.temp_bytes, st);
    if (rc) return rc;
    APRB_TIMED("sub_flags_kernel", st, 1, (sub_flags_kernel<KeyT><<<cdiv(N + 1, T), T, 0, st>>>(keys_out, N, w.flags)));
    rc = exclusive_scan_i32(w.flags, w.pos, N + 1, w.temp, w.temp_bytes, st);
    if (rc) return rc;
    APRB_TIMED("sub_lens_kernel", st, 1, (sub_lens_kernel<<<1, 256, 0, st>>>(w.pos, w.off, B, 0, d_out_lens, w.out_start, d_out_M)));
    APRB_TIMED("o3d_emit_kernel", st, 1, (o3d_emit_kernel<KeyT><<<cdiv(N, T), T, 0, st>>>(d_raw, stride, N, keys_out, w.vals_out, w.pos, w.off, B, w.out_start, d_out_pts)));
    APRB_LAUNCH_OK();
    return APRB_OK;
}

extern "C" size_t aprb_voxel_downsample_ws_bytes(int N, int B) {
    if (N < 0 || B < 0) return 0;
    return aprb_grid_subsample_ws_bytes(N, B, 0) + align256(sizeof(O3dGrid) * (size_t)(B > 0 ? B : 1)) + 256;
}

extern "C" int aprb_voxel_downsample_raw(const float* d_raw, int stride, const int32_t* d_lens, int B, int N, double voxel_size,
                                         float* d_out_pts, int32_t* d_out_lens, int32_t* d_out_M, int32_t* d_status,
                                         int key_bits, void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(B >= 1 && N >= 0 && stride >= 3, "need B >= 1, N >= 0 and a row stride of at least 3 floats");
    APRB_REQUIRE(d_lens && d_out_lens && d_out_M, "null length/output pointer");
    APRB_REQUIRE(voxel_size > 0.0, "voxel_size must be positive");
    APRB_REQUIRE(key_bits == 32 || key_bits == 64, "key_bits must be 32 or 64");
    if (N == 0) {
        if (d_status) APRB_CUDA_OK(cudaMemsetAsync(d_status, 0, sizeof(int), st));
        APRB_CUDA_OK(cudaMemsetAsync(d_out_lens, 0, sizeof(int) * B, st));
        APRB_CUDA_OK(cudaMemsetAsync(d_out_M, 0, sizeof(int), st));
        return APRB_OK;
    }
    APRB_REQUIRE(d_raw && d_out_pts && d_ws, "null point/workspace pointer");
    Carver c(d_ws, ws_bytes);
    SubWs w;
    carve_sub(c, N, B, &w);
    O3dGrid* grids = c.take<O3dGrid>(B);
    if (!c.ok()) { set_error("aprb_voxel_downsample_raw: workspace too small (%zu < %zu)", ws_bytes, c.off); return APRB_ERR_WORKSPACE; }
    if (key_bits == 32) return run_voxel_raw<uint32_t>(d_raw, stride, d_lens, B, N, voxel_size, d_out_pts, d_out_lens, d_out_M, d_status, w, grids, st);
    return run_voxel_raw<uint64_t>(d_raw, stride, d_lens, B, N, voxel_size, d_out_pts, d_out_lens, d_out_M, d_status, w, grids, st);
}

static int subsample_impl(const float* d_pts, const int32_t* d_lens, int B, int N, float dl, int max_p,
                          const float* d_feats, int fdim, const int32_t* d_classes, int ldim, float* d_out_pts, int32_t* d_out_lens,
                          int32_t* d_out_M, float* d_out_feats, int32_t* d_out_classes, int32_t* d_status, int key_bits,
                          void* d_ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    APRB_REQUIRE(!(d_classes && (ldim <= 0 || !d_out_classes)), "classes given without ldim/out buffer");
    APRB_REQUIRE(B >= 1 && N >= 0, "need B >= 1 and N >= 0");
    APRB_REQUIRE(d_lens && d_out_lens && d_out_M, "null length/output pointer");
    APRB_REQUIRE(dl > 0.f, "sampleDl must be positive");
    APRB_REQUIRE(key_bits == 32 || key_bits == 64, "key_bits must be 32 or 64");
    APRB_REQUIRE(!(d_feats && (fdim <= 0 || !d_out_feats)), "features given without fdim/out buffer");
    if (N == 0) {
        if (d_status) APRB_CUDA_OK(cudaMemsetAsync(d_status, 0, sizeof(int), st));
        APRB_CUDA_OK(cudaMemsetAsync(d_out_lens, 0, sizeof(int) * B, st));
        APRB_CUDA_OK(cudaMemsetAsync(d_out_M, 0, sizeof(int), st));
        return APRB_OK;
    }
    APRB_REQUIRE(d_pts && d_out_pts && d_ws, "null point/workspace pointer");
    Carver c(d_ws, ws_bytes);
    SubWs w;
    carve_sub(c, N, B, &w);
    if (!c.ok()) { set_error("aprb_grid_subsample_batch: workspace too small (%zu < %zu)", ws_bytes, c.off); return APRB_ERR_WORKSPACE; }
    if (key_bits == 32)
        return run_subsample<uint32_t>(d_pts, d_lens, B, N, dl, max_p, d_feats, fdim, d_out_pts, d_out_lens, d_out_M, d_out_feats, d_status, w, st, d_classes, ldim, d_out_classes);
    return run_subsample<uint64_t>(d_pts, d_lens, B, N, dl, max_p, d_feats, fdim, d_out_pts, d_out_lens, d_out_M, d_out_feats, d_status, w, st, d_classes, ldim, d_out_classes);
}

extern "C" int aprb_grid_subsample_batch(const float* d_pts, const int32_t* d_lens, int B, int N, float dl, int max_p,
                                         const float* d_feats, int fdim, float* d_out_pts, int32_t* d_out_lens,
                                         int32_t* d_out_M, float* d_out_feats, int32_t* d_status, int key_bits,
                                         void* d_ws, size_t ws_bytes, void* stream) {
    return subsample_impl(d_pts, d_lens, B, N, dl, max_p, d_feats, fdim, nullptr, 0, d_out_pts, d_out_lens, d_out_M, d_out_feats,
                          nullptr, d_status, key_bits, d_ws, ws_bytes, stream);
}

extern "C" int aprb_grid_subsample_batch_labels(const float* d_pts, const int32_t* d_lens, int B, int N, float dl, int max_p,
                                                const float* d_feats, int fdim, const int32_t* d_classes, int ldim,
                                                float* d_out_pts, int32_t* d_out_lens, int32_t* d_out_M, float* d_out_feats,
                                                int32_t* d_out_classes, int32_t* d_status, int key_bits,
                                                void* d_ws, size_t ws_bytes, void* stream) {
    return subsample_impl(d_pts, d_lens, B, N, dl, max_p, d_feats, fdim, d_classes, ldim, d_out_pts, d_out_lens, d_out_M, d_out_feats,
                          d_out_classes, d_status, key_bits, d_ws, ws_bytes, stream);
}
