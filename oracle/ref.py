"""TEST INFRASTRUCTURE ONLY — Python face of the CPU oracles.

* `Oracle`  : our plain-C restatement (oracle/_build/liboracle.so, source oracle/oracle_l1.c).
* `RefL1`   : the reference's own object code (oracle/_ref/libapr_ref.so, built by oracle/Makefile from
              /root/reference/Predator_APR/cpp_wrappers/{cpp_utils/cloud/cloud.cpp,
              cpp_neighbors/neighbors/neighbors.cpp, cpp_subsampling/grid_subsampling/grid_subsampling.cpp}).
* `collate_ref`: restatement of the pyramid schedule datasets/dataloader.py:72-198 (cannot be imported: open3d
              at dataloader.py:1), parameterised on the two native callables.
* `calibrate_ref`: restatement of datasets/dataloader.py:200-232.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product package apr_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)


def build(verbose=False):
    """Compile oracle/_build/liboracle.so and, when /root/reference is present, oracle/_ref/libapr_ref.so."""
    r = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)


def _f32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(np.asarray(a), dtype=np.int32)


class Oracle:
    """ctypes binding of oracle_l1.c."""

    def __init__(self):
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        self.lib.orc_grid_subsample_batch.restype = C.c_int
        self.lib.orc_grid_subsample_batch.argtypes = [_f32p, C.c_int, _i32p, C.c_int, C.c_float, C.c_int, _f32p, _i32p]
        self.lib.orc_radius_neighbors_batch.restype = C.c_int
        self.lib.orc_radius_neighbors_batch.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _i32p, _i32p, C.c_int,
                                                        C.c_float, C.c_int, C.POINTER(_i32p), _i32p]
        self.lib.orc_free.argtypes = [C.c_void_p]

    def subsample_batch(self, points, batches, sampleDl=0.1, max_p=0):
        p, l = _f32(points), _i32(batches)
        out = np.empty((max(len(p), 1), 3), np.float32)
        ol = np.empty(len(l), np.int32)
        m = self.lib.orc_grid_subsample_batch(p.ctypes.data_as(_f32p), len(p), l.ctypes.data_as(_i32p), len(l),
                                              np.float32(sampleDl), int(max_p), out.ctypes.data_as(_f32p),
                                              ol.ctypes.data_as(_i32p))
        return out[:m].copy(), ol

    def batch_query(self, queries, supports, q_batches, s_batches, radius=0.1, max_neighbors=0, return_counts=False):
        q, s, ql, sl = _f32(queries), _f32(supports), _i32(q_batches), _i32(s_batches)
        ptr = _i32p()
        counts = np.zeros(max(len(q), 1), np.int32)
        w = self.lib.orc_radius_neighbors_batch(q.ctypes.data_as(_f32p), len(q), s.ctypes.data_as(_f32p), len(s),
                                                ql.ctypes.data_as(_i32p), sl.ctypes.data_as(_i32p), len(ql),
                                                np.float32(radius), int(max_neighbors), C.byref(ptr),
                                                counts.ctypes.data_as(_i32p))
        out = np.ctypeslib.as_array(ptr, shape=(len(q) * w + 1,))[:len(q) * w].reshape(len(q), w).copy()
        self.lib.orc_free(ptr)
        return (out, counts[:len(q)]) if return_counts else out


class RefL1:
    """ctypes binding of the reference's own object code (oracle/_ref)."""

    def __init__(self):
        path = os.path.join(_HERE, "_ref", "libapr_ref.so")
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref/libapr_ref.so missing and /root/reference absent")
        self.lib = C.CDLL(path)
        self.lib.ref_batch_grid_subsampling.restype = C.c_int
        self.lib.ref_batch_grid_subsampling.argtypes = [_f32p, C.c_int, _i32p, C.c_int, C.c_float, C.c_int,
                                                        C.POINTER(_f32p), _i32p]
        if hasattr(self.lib, "ref_batch_grid_subsampling_full"):
            self.lib.ref_batch_grid_subsampling_full.restype = C.c_int
            self.lib.ref_batch_grid_subsampling_full.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _i32p, C.c_int, _i32p, C.c_int,
                                                                 C.c_float, C.c_int, C.POINTER(_f32p), C.POINTER(_f32p),
                                                                 C.POINTER(_i32p), _i32p]
        self.lib.ref_batch_neighbors.restype = C.c_int
        self.lib.ref_batch_neighbors.argtypes = [_f32p, C.c_int, _f32p, C.c_int, _i32p, _i32p, C.c_int, C.c_float,
                                                 C.c_int, C.POINTER(_i32p)]
        self.lib.ref_free.argtypes = [C.c_void_p]

    @staticmethod
    def available():
        return os.path.exists(os.path.join(_HERE, "_ref", "libapr_ref.so"))

    def subsample_batch(self, points, batches, sampleDl=0.1, max_p=0):
        p, l = _f32(points), _i32(batches)
        ptr = _f32p()
        ol = np.empty(len(l), np.int32)
        m = self.lib.ref_batch_grid_subsampling(p.ctypes.data_as(_f32p), len(p), l.ctypes.data_as(_i32p), len(l),
                                                np.float32(sampleDl), int(max_p), C.byref(ptr), ol.ctypes.data_as(_i32p))
        out = np.ctypeslib.as_array(ptr, shape=(max(m, 1) * 3,))[:m * 3].reshape(m, 3).copy()
        self.lib.ref_free(ptr)
        return out, ol

    def subsample_batch_full(self, points, batches, features=None, classes=None, sampleDl=0.1, max_p=0):
        """batch_grid_subsampling with features / labels: (points, lens, feats or None, classes [M,ldim] or None), rows in
        the reference's own order. Labels: ldim == 1 or a single cloud only (see ref_shim.cpp)."""
        p, l = _f32(points), _i32(batches)
        f = _f32(features) if features is not None else None
        c = _i32(classes).reshape(len(p), -1) if classes is not None else None
        fdim, ldim = (f.shape[1] if f is not None else 0), (c.shape[1] if c is not None else 0)
        assert ldim <= 1 or len(l) == 1
        pp, pf, pc = _f32p(), _f32p(), _i32p()
        ol = np.empty(len(l), np.int32)
        m = self.lib.ref_batch_grid_subsampling_full(
            p.ctypes.data_as(_f32p), len(p), f.ctypes.data_as(_f32p) if f is not None else None, fdim,
            c.ctypes.data_as(_i32p) if c is not None else None, ldim, l.ctypes.data_as(_i32p), len(l),
            np.float32(sampleDl), int(max_p), C.byref(pp), C.byref(pf), C.byref(pc), ol.ctypes.data_as(_i32p))
        out = np.ctypeslib.as_array(pp, shape=(max(m, 1) * 3,))[:m * 3].reshape(m, 3).copy()
        of = np.ctypeslib.as_array(pf, shape=(max(m * fdim, 1),))[:m * fdim].reshape(m, fdim).copy() if fdim else None
        oc = np.ctypeslib.as_array(pc, shape=(max(m * ldim, 1),))[:m * ldim].reshape(m, ldim).copy() if ldim else None
        for q in (pp, pf, pc):
            self.lib.ref_free(q)
        return out, ol, of, oc

    def batch_query(self, queries, supports, q_batches, s_batches, radius=0.1, variant="nanoflann"):
        q, s, ql, sl = _f32(queries), _f32(supports), _i32(q_batches), _i32(s_batches)
        ptr = _i32p()
        w = self.lib.ref_batch_neighbors(q.ctypes.data_as(_f32p), len(q), s.ctypes.data_as(_f32p), len(s),
                                         ql.ctypes.data_as(_i32p), sl.ctypes.data_as(_i32p), len(ql),
                                         np.float32(radius), 0 if variant == "nanoflann" else 1, C.byref(ptr))
        out = np.ctypeslib.as_array(ptr, shape=(len(q) * w + 1,))[:len(q) * w].reshape(len(q), w).copy()
        self.lib.ref_free(ptr)
        return out


# ----------------------------------------------------------------------------------------------------------------
# canonicalisers / comparison helpers
# ----------------------------------------------------------------------------------------------------------------

def lexsort_rows_per_cloud(points, lens):
    """Sort rows of each cloud lexicographically (x, then y, then z) — order-free comparison of subsample outputs."""
    out, off = [], 0
    for n in lens:
        blk = points[off:off + n]
        out.append(blk[np.lexsort((blk[:, 2], blk[:, 1], blk[:, 0]))])
        off += n
    return np.concatenate(out, 0) if out else points


def d2_rows(queries, supports, idx):
    """fp32 squared distances computed exactly like nanoflann.hpp:432-440 (x,y,z order, no FMA); pads -> +inf."""
    q = np.asarray(queries, np.float32)
    s = np.concatenate([np.asarray(supports, np.float32), np.full((1, 3), np.float32(np.inf), np.float32)], 0)
    d = q[:, None, :] - s[idx]
    with np.errstate(invalid="ignore"):
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    d2[idx == len(supports)] = np.inf
    return d2.astype(np.float32)


def equal_modulo_ties(a, b, queries, supports):
    """True iff index matrices a and b agree except for permutations inside runs of equal fp32 d2.
    Returns (ok, n_rows_differing, n_nontie_mismatches)."""
    if a.shape != b.shape:
        return False, -1, -1
    diff = a != b
    if not diff.any():
        return True, 0, 0
    da, db = d2_rows(queries, supports, a), d2_rows(queries, supports, b)
    nontie = int((diff & (da != db)).sum())
    rows = np.where(diff.any(1))[0]
    for r in rows:   # within each row the multiset of indices must agree
        if not np.array_equal(np.sort(a[r]), np.sort(b[r])):
            nontie += 1
    return nontie == 0, len(rows), nontie


# ----------------------------------------------------------------------------------------------------------------
# pyramid schedule (datasets/dataloader.py:72-198) and neighbourhood calibration (:200-232)
# ----------------------------------------------------------------------------------------------------------------

def collate_ref(points, lengths, config, neighborhood_limits, subsample_batch, batch_query):
    """points fp32 [N,3] stacked pair, lengths int32 [B]. `subsample_batch(points, lens, sampleDl=)` and
    `batch_query(q, s, ql, sl, radius=)` are the two native callables (reference or oracle).
    Returns dict(points, neighbors, pools, upsamples, stack_lengths) of numpy arrays (int32 indices; the reference
    widens them to int64 at dataloader.py:164-166)."""
    def neigh(q, s, ql, sl, r, lim):                                   # dataloader.py:55-70
        nb = batch_query(q, s, ql, sl, radius=r)
        return nb[:, :lim] if lim > 0 else nb

    r_normal = config.first_subsampling_dl * config.conv_radius        # :93
    pts, lens = np.asarray(points, np.float32), np.asarray(lengths, np.int32)
    layer_blocks, layer = [], 0
    out = dict(points=[], neighbors=[], pools=[], upsamples=[], stack_lengths=[])
    arch = config.architecture
    for block_i, block in enumerate(arch):
        if 'global' in block or 'upsample' in block:                   # :107
            break
        if not ('pool' in block or 'strided' in block):                # :111-114
            layer_blocks += [block]
            if block_i < len(arch) - 1 and not ('upsample' in arch[block_i + 1]):
                continue
        if layer_blocks:                                               # :119-125
            if any('deformable' in blck for blck in layer_blocks[:-1]):
                r = r_normal * config.deform_radius / config.conv_radius
            else:
                r = r_normal
            conv_i = neigh(pts, pts, lens, lens, r, neighborhood_limits[layer])
        else:
            conv_i = np.zeros((0, 1), np.int32)
        if 'pool' in block or 'strided' in block:                      # :135-153
            dl = 2 * r_normal / config.conv_radius
            pool_p, pool_b = subsample_batch(pts, lens, sampleDl=dl)
            r = r_normal * config.deform_radius / config.conv_radius if 'deformable' in block else r_normal
            pool_i = neigh(pool_p, pts, pool_b, lens, r, neighborhood_limits[layer])
            up_i = neigh(pts, pool_p, lens, pool_b, 2 * r, neighborhood_limits[layer])
        else:                                                          # :155-160
            pool_i = np.zeros((0, 1), np.int32)
            pool_p = np.zeros((0, 3), np.float32)
            pool_b = np.zeros((0,), np.int32)
            up_i = np.zeros((0, 1), np.int32)
        out['points'].append(pts); out['neighbors'].append(conv_i); out['pools'].append(pool_i)
        out['upsamples'].append(up_i); out['stack_lengths'].append(lens)
        pts, lens = pool_p, pool_b                                     # :170-171
        r_normal *= 2; layer += 1; layer_blocks = []                   # :174-176
    return out


def calibrate_ref(pairs, config, subsample_batch, batch_query, keep_ratio=0.8, samples_threshold=2000):
    """pairs: iterable of (points [N,3], lengths [2]). Restates datasets/dataloader.py:200-232."""
    hist_n = int(np.ceil(4 / 3 * np.pi * (config.deform_radius + 1) ** 3))      # :205 -> 905
    hists = np.zeros((config.num_layers, hist_n), np.int32)
    for pts, lens in pairs:
        b = collate_ref(pts, lens, config, [hist_n] * 5, subsample_batch, batch_query)
        counts = [np.sum(nb < nb.shape[0], axis=1) for nb in b['neighbors']]    # :214
        hists += np.vstack([np.bincount(c, minlength=hist_n)[:hist_n] for c in counts])
        if np.min(np.sum(hists, axis=1)) > samples_threshold:                   # :223
            break
    cumsum = np.cumsum(hists.T, axis=0)
    return np.sum(cumsum < (keep_ratio * cumsum[hist_n - 1, :]), axis=0)        # :226-227
