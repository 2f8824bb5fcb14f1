/* TEST INFRASTRUCTURE ONLY — CPU oracle for the L1 half of the hot path (grid subsampling + radius search).
 *
 * Plain-C restatement (written from scratch, not copied) of the reference algorithms, all paths relative to
 * /root/reference/Predator_APR/cpp_wrappers:
 *   orc_grid_subsample_batch   <- cpp_subsampling/grid_subsampling/grid_subsampling.cpp:5-106 (per cloud) and
 *                                 :109-211 (batch loop, max_p truncation); min/max: cpp_utils/cloud/cloud.cpp:27-67;
 *                                 accumulate: grid_subsampling.h:74-79; barycentre: grid_subsampling.cpp:87 with
 *                                 operator*(PointXYZ,float) cpp_utils/cloud/cloud.h:120-123.
 *   orc_radius_neighbors_batch <- cpp_neighbors/neighbors/neighbors.cpp:125-208 (batch_ordered_neighbors: ascending
 *                                 d2, ties by ascending support index) which is the canonical form of the live
 *                                 neighbors.cpp:211-332 (batch_nanoflann_neighbors); distance arithmetic
 *                                 cpp_utils/nanoflann/nanoflann.hpp:432-440, membership :249-253 (strict <),
 *                                 truncation datasets/dataloader.py:66-70.
 * Pinned against the reference's own object code (oracle/_ref/libapr_ref.so) by tests/test_oracle.py:
 * subsample rows bit-identical as a set, neighbour matrices bit-identical to batch_ordered_neighbors and identical
 * to batch_nanoflann_neighbors modulo permutation inside equal-d2 runs.
 *
 * Deviation (documented in DESIGN.md): subsample rows are emitted in ascending voxel-key order per cloud (the
 * reference emits libstdc++ unordered_map iteration order, grid_subsampling.cpp:85, which is unspecified).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this. Build: -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t key; int32_t count; float sx, sy, sz; } voxel_t;

static int cmp_voxel(const void* a, const void* b) {
    uint64_t ka = ((const voxel_t*)a)->key, kb = ((const voxel_t*)b)->key;
    return ka < kb ? -1 : (ka > kb ? 1 : 0);
}

/* float -> size_t the way x86-64 gcc does it for values below 2^63 (cvttss2si): negatives wrap. */
static uint64_t f2sz(float f) { return (uint64_t)(int64_t)f; }

/* One cloud. Returns number of voxels; writes barycentres in ascending key order to out (capacity n rows). */
static int grid_subsample_one(const float* p, int n, float dl, float* out) {
    if (n <= 0) return 0;
    float mnx = p[0], mny = p[1], mnz = p[2], mxx = p[0], mxy = p[1], mxz = p[2];
    for (int i = 0; i < n; ++i) {                       /* cloud.cpp:27-67 */
        const float* q = p + 3 * i;
        if (q[0] < mnx) mnx = q[0]; if (q[1] < mny) mny = q[1]; if (q[2] < mnz) mnz = q[2];
        if (q[0] > mxx) mxx = q[0]; if (q[1] > mxy) mxy = q[1]; if (q[2] > mxz) mxz = q[2];
    }
    float inv = 1 / dl;                                  /* grid_subsampling.cpp:27: (1/sampleDl) in fp32 */
    float ox = floorf(mnx * inv) * dl, oy = floorf(mny * inv) * dl, oz = floorf(mnz * inv) * dl;
    uint64_t NX = f2sz(floorf((mxx - ox) / dl)) + 1;     /* :30-31 */
    uint64_t NY = f2sz(floorf((mxy - oy) / dl)) + 1;
    (void)mxz;
    /* open-addressing table keyed by voxel key; accumulation in input order (:50-77) */
    size_t cap = 16; while (cap < (size_t)n * 2) cap <<= 1;
    voxel_t* tab = (voxel_t*)calloc(cap, sizeof(voxel_t));
    int m = 0;
    for (int i = 0; i < n; ++i) {
        const float* q = p + 3 * i;
        uint64_t iX = f2sz(floorf((q[0] - ox) / dl));
        uint64_t iY = f2sz(floorf((q[1] - oy) / dl));
        uint64_t iZ = f2sz(floorf((q[2] - oz) / dl));
        uint64_t key = iX + NX * iY + NX * NY * iZ;
        size_t h = (size_t)((key * 0x9E3779B97F4A7C15ull) >> 20) & (cap - 1);
        while (tab[h].count && tab[h].key != key) h = (h + 1) & (cap - 1);
        if (!tab[h].count) { tab[h].key = key; ++m; }
        tab[h].count += 1;                               /* grid_subsampling.h:74-79 */
        tab[h].sx += q[0]; tab[h].sy += q[1]; tab[h].sz += q[2];
    }
    voxel_t* vox = (voxel_t*)malloc(sizeof(voxel_t) * (size_t)m);
    int j = 0;
    for (size_t h = 0; h < cap; ++h) if (tab[h].count) vox[j++] = tab[h];
    qsort(vox, (size_t)m, sizeof(voxel_t), cmp_voxel);   /* canonical order (deviation, see header) */
    for (j = 0; j < m; ++j) {
        float w = (float)(1.0 / vox[j].count);           /* :87 double reciprocal narrowed by operator*(P,float) */
        out[3 * j + 0] = vox[j].sx * w; out[3 * j + 1] = vox[j].sy * w; out[3 * j + 2] = vox[j].sz * w;
    }
    free(vox); free(tab);
    return m;
}

/* Stacked batch. out_pts has capacity N rows. Returns M = total rows. (grid_subsampling.cpp:109-211) */
int orc_grid_subsample_batch(const float* pts, int N, const int32_t* lens, int B, float dl, int max_p,
                             float* out_pts, int32_t* out_lens) {
    if (max_p < 1) max_p = N;
    int off = 0, M = 0;
    float* tmp = (float*)malloc(sizeof(float) * 3 * (size_t)(N > 0 ? N : 1));
    for (int b = 0; b < B; ++b) {
        int m = grid_subsample_one(pts + 3 * (size_t)off, lens[b], dl, tmp);
        if (m > max_p) m = max_p;                        /* :181-204 keep the first max_p (canonical order here) */
        memcpy(out_pts + 3 * (size_t)M, tmp, sizeof(float) * 3 * (size_t)m);
        out_lens[b] = m; M += m; off += lens[b];
    }
    free(tmp);
    return M;
}

typedef struct { float d2; int32_t idx; } cand_t;
static int cmp_cand(const void* a, const void* b) {
    const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
    if (x->d2 < y->d2) return -1; if (x->d2 > y->d2) return 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}

/* Brute-force radius search in the canonical (d2, index) order. *out_idx is malloc'ed [Nq, width] with
 * width = max_count (or min(max_count, max_neighbors) when max_neighbors > 0); counts (optional) receives the
 * untruncated per-query neighbour count. Pads = Ns (neighbors.cpp:322-324). Returns width. */
int orc_radius_neighbors_batch(const float* q, int Nq, const float* s, int Ns, const int32_t* ql, const int32_t* sl,
                               int B, float radius, int max_neighbors, int32_t** out_idx, int32_t* counts) {
    float r2 = radius * radius;                          /* neighbors.cpp:226 */
    cand_t** rows = (cand_t**)calloc((size_t)(Nq > 0 ? Nq : 1), sizeof(cand_t*));
    int32_t* cnt = (int32_t*)calloc((size_t)(Nq > 0 ? Nq : 1), sizeof(int32_t));
    int max_count = 0, qoff = 0, soff = 0;
    cand_t* buf = (cand_t*)malloc(sizeof(cand_t) * (size_t)(Ns > 0 ? Ns : 1));
    for (int b = 0; b < B; ++b) {
        for (int i = qoff; i < qoff + ql[b]; ++i) {
            float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
            int c = 0;
            for (int j = soff; j < soff + sl[b]; ++j) {
                float dx = qx - s[3 * j], dy = qy - s[3 * j + 1], dz = qz - s[3 * j + 2];
                float d2 = dx * dx + dy * dy + dz * dz;  /* nanoflann.hpp:432-440 order, no FMA */
                if (d2 < r2) { buf[c].d2 = d2; buf[c].idx = j; ++c; }   /* :249-253 strict */
            }
            qsort(buf, (size_t)c, sizeof(cand_t), cmp_cand);
            rows[i] = (cand_t*)malloc(sizeof(cand_t) * (size_t)(c > 0 ? c : 1));
            memcpy(rows[i], buf, sizeof(cand_t) * (size_t)c);
            cnt[i] = c; if (c > max_count) max_count = c;
        }
        qoff += ql[b]; soff += sl[b];
    }
    int w = max_count;
    if (max_neighbors > 0 && w > max_neighbors) w = max_neighbors;   /* dataloader.py:66-70 */
    int32_t* out = (int32_t*)malloc(sizeof(int32_t) * ((size_t)Nq * (size_t)w + 1));
    for (int i = 0; i < Nq; ++i) {
        for (int j = 0; j < w; ++j) out[(size_t)i * w + j] = j < cnt[i] ? rows[i][j].idx : Ns;
        if (counts) counts[i] = cnt[i];
        free(rows[i]);
    }
    free(rows); free(cnt); free(buf);
    *out_idx = out;
    return w;
}

/* d2 exactly as the search computes it, for tie analysis in tests. */
float orc_d2(const float* a, const float* b) {
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return dx * dx + dy * dy + dz * dz;
}

void orc_free(void* p) { free(p); }
