// TEST INFRASTRUCTURE ONLY (oracle/_ref): C-ABI shim over the UNMODIFIED reference C++ core.
//
// This file is ours; it #includes the reference's own headers where they lie under
// /root/reference (never copied into the repo) and exposes the three live entry points as
// extern "C" so the tests / bench cpu_baseline leg can call the reference's object code via ctypes:
//   batch_grid_subsampling     Predator_APR/cpp_wrappers/cpp_subsampling/grid_subsampling/grid_subsampling.cpp:109-211
//   batch_nanoflann_neighbors  Predator_APR/cpp_wrappers/cpp_neighbors/neighbors/neighbors.cpp:211-332
//   batch_ordered_neighbors    Predator_APR/cpp_wrappers/cpp_neighbors/neighbors/neighbors.cpp:125-208
// The reference's own CPython wrappers do not compile against numpy 2.x, hence the shim.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load the result.
#include "cpp_subsampling/grid_subsampling/grid_subsampling.h"
#include "cpp_neighbors/neighbors/neighbors.h"
#include <cstring>
#include <cstdlib>

extern "C" {

// Returns M (number of subsampled points); *out_pts is malloc'ed [M,3] (free with ref_free).
int ref_batch_grid_subsampling(const float* pts, int N, const int* lens, int B, float dl, int max_p,
                               float** out_pts, int* out_lens) {
    std::vector<PointXYZ> op((const PointXYZ*)pts, (const PointXYZ*)pts + N), sp;
    std::vector<float> of, sf;
    std::vector<int> oc, sc, ob(lens, lens + B), sb;
    batch_grid_subsampling(op, sp, of, sf, oc, sc, ob, sb, dl, max_p);
    int M = (int)sp.size();
    *out_pts = (float*)malloc(sizeof(float) * 3 * (M > 0 ? M : 1));
    if (M) memcpy(*out_pts, sp.data(), sizeof(float) * 3 * M);
    for (int b = 0; b < B; ++b) out_lens[b] = sb[b];
    return M;
}

// The same with features [N,fdim] and labels [N,ldim] (either may be NULL / 0): outputs malloc'ed, rows in the reference's
// own (unordered_map) order. The reference slices the labels of cloud b as [sum_b*ldim, sum_b + len*ldim)
// (grid_subsampling.cpp:167-168), which is only right for ldim == 1 or B == 1: callers keep to those.
int ref_batch_grid_subsampling_full(const float* pts, int N, const float* feats, int fdim, const int* classes, int ldim,
                                    const int* lens, int B, float dl, int max_p, float** out_pts, float** out_feats,
                                    int** out_classes, int* out_lens) {
    std::vector<PointXYZ> op((const PointXYZ*)pts, (const PointXYZ*)pts + N), sp;
    std::vector<float> of, sf;
    if (feats && fdim > 0) of.assign(feats, feats + (size_t)N * fdim);
    std::vector<int> oc, sc, ob(lens, lens + B), sb;
    if (classes && ldim > 0) oc.assign(classes, classes + (size_t)N * ldim);
    batch_grid_subsampling(op, sp, of, sf, oc, sc, ob, sb, dl, max_p);
    int M = (int)sp.size();
    *out_pts = (float*)malloc(sizeof(float) * 3 * (M > 0 ? M : 1));
    if (M) memcpy(*out_pts, sp.data(), sizeof(float) * 3 * M);
    *out_feats = (float*)malloc(sizeof(float) * (sf.size() ? sf.size() : 1));
    if (sf.size()) memcpy(*out_feats, sf.data(), sizeof(float) * sf.size());
    *out_classes = (int*)malloc(sizeof(int) * (sc.size() ? sc.size() : 1));
    if (sc.size()) memcpy(*out_classes, sc.data(), sizeof(int) * sc.size());
    for (int b = 0; b < B; ++b) out_lens[b] = sb[b];
    return M;
}

// variant: 0 = batch_nanoflann_neighbors (the live one), 1 = batch_ordered_neighbors.
// Returns width (max_count); *out_idx is malloc'ed [Nq,width].
int ref_batch_neighbors(const float* q, int Nq, const float* s, int Ns, const int* ql, const int* sl, int B,
                        float radius, int variant, int** out_idx) {
    std::vector<PointXYZ> qv((const PointXYZ*)q, (const PointXYZ*)q + Nq), sv((const PointXYZ*)s, (const PointXYZ*)s + Ns);
    std::vector<int> qb(ql, ql + B), sb(sl, sl + B), res;
    if (variant == 0) batch_nanoflann_neighbors(qv, sv, qb, sb, res, radius);
    else batch_ordered_neighbors(qv, sv, qb, sb, res, radius);
    int w = Nq > 0 ? (int)(res.size() / (size_t)Nq) : 0;
    *out_idx = (int*)malloc(sizeof(int) * (res.size() ? res.size() : 1));
    if (res.size()) memcpy(*out_idx, res.data(), sizeof(int) * res.size());
    return w;
}

void ref_free(void* p) { free(p); }

}  // extern "C"
