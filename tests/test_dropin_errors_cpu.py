"""Host logic of the drop-in modules without a GPU: the argument checks of the reference's CPython wrappers
(/root/reference/Predator_APR/cpp_wrappers/cpp_neighbors/wrapper.cpp:71-75,127-171,201-205 and
cpp_subsampling/wrapper.cpp:75-96,154-226,267-271) run before any device work, raise RuntimeError with the reference's
strings, keep the keyword-only split of the "$" format, and a call that passes them fails loudly on a CPU-only box
(no CPU fallback)."""
import numpy as np
import pytest
import torch

import cpp_wrappers.cpp_neighbors.radius_neighbors as cpp_neighbors
import cpp_wrappers.cpp_subsampling.grid_subsampling as cpp_subsampling


def _pts(n, seed=0):
    return np.random.default_rng(seed).normal(size=(n, 3)).astype(np.float32)


def test_import_paths_are_the_reference_ones():
    # datasets/dataloader.py:5-6 imports exactly these two module paths
    assert callable(cpp_neighbors.batch_query)
    assert callable(cpp_subsampling.subsample_batch) and callable(cpp_subsampling.subsample)


def test_batch_query_argument_errors_match_reference_strings():
    p, l = _pts(10), np.array([10], np.int32)
    with pytest.raises(RuntimeError, match=r"Wrong dimensions : query.shape is not \(N, 3\)"):
        cpp_neighbors.batch_query(np.zeros((10, 2), np.float32), p, l, l, radius=1.0)
    with pytest.raises(RuntimeError, match=r"Wrong dimensions : support.shape is not \(N, 3\)"):
        cpp_neighbors.batch_query(p, np.zeros((10,), np.float32), l, l, radius=1.0)
    with pytest.raises(RuntimeError, match=r"Wrong dimensions : queries_batches.shape is not \(B,\)"):
        cpp_neighbors.batch_query(p, p, np.zeros((1, 1), np.int32), l, radius=1.0)
    with pytest.raises(RuntimeError, match="Wrong number of batch elements: different for queries and supports"):
        cpp_neighbors.batch_query(p, p, np.array([4, 6], np.int32), l, radius=1.0)
    with pytest.raises(RuntimeError, match="Error converting query points to numpy arrays of type float32"):
        cpp_neighbors.batch_query([["a", "b", "c"]], p, l, l, radius=1.0)
    with pytest.raises(RuntimeError, match="^Error$"):                  # empty result (wrapper.cpp:201-205)
        cpp_neighbors.batch_query(np.zeros((0, 3), np.float32), p, np.array([0], np.int32), l, radius=1.0)
    with pytest.raises(TypeError):                                      # options after "$" are keyword-only
        cpp_neighbors.batch_query(p, p, l, l, 1.0)


def test_subsample_argument_errors_match_reference_strings():
    p, l = _pts(10), np.array([10], np.int32)
    with pytest.raises(RuntimeError, match='Error parsing method. Valid method names are "barycenters" and "voxelcenters"'):
        cpp_subsampling.subsample_batch(p, l, sampleDl=0.3, method="median")
    with pytest.raises(RuntimeError, match=r"Wrong dimensions : points.shape is not \(N, 3\)"):
        cpp_subsampling.subsample_batch(np.zeros((10, 4), np.float32), l, sampleDl=0.3)
    with pytest.raises(RuntimeError, match=r"Wrong dimensions : batches.shape is not \(B,\)"):
        cpp_subsampling.subsample_batch(p, np.zeros((1, 1), np.int32), sampleDl=0.3)
    with pytest.raises(RuntimeError, match=r"Wrong dimensions : features.shape is not \(N, d\)"):
        cpp_subsampling.subsample_batch(p, l, features=np.zeros((9, 2), np.float32), sampleDl=0.3)
    with pytest.raises(RuntimeError, match="Error converting input points to numpy arrays of type float32"):
        cpp_subsampling.subsample_batch([["x", "y", "z"]], l, sampleDl=0.3)
    with pytest.raises(RuntimeError, match="^Error$"):                  # empty result (wrapper.cpp:267-271)
        cpp_subsampling.subsample_batch(np.zeros((0, 3), np.float32), np.array([0], np.int32), sampleDl=0.3)
    with pytest.raises(RuntimeError, match=r"Wrong dimensions : classes.shape is not \(N,\) or \(N, d\)"):
        cpp_subsampling.subsample_batch(p, l, classes=np.zeros((10, 1, 1), np.int32), sampleDl=0.3)
    with pytest.raises(RuntimeError, match="Error converting input classes to numpy arrays of type int32"):
        cpp_subsampling.subsample_batch(p, l, classes=[["a"]] * 10, sampleDl=0.3)
    with pytest.raises(TypeError):                                      # keyword-only options
        cpp_subsampling.subsample_batch(p, l, 0.3)
    with pytest.raises(RuntimeError, match=r"Wrong dimensions : points.shape is not \(N, 3\)"):
        cpp_subsampling.subsample(np.zeros((10,), np.float32), sampleDl=0.3)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour of a box without a GPU")
def test_valid_call_without_gpu_fails_loudly_not_silently():
    """Inputs that pass the checks need the device: on a CPU-only box the call must raise, never compute on the CPU."""
    p, l = _pts(64), np.array([64], np.int32)
    with pytest.raises(Exception):
        cpp_neighbors.batch_query(p, p, l, l, radius=1.0)
    with pytest.raises(Exception):
        cpp_subsampling.subsample_batch(p, l, sampleDl=0.3)
    from apr_b200 import _native, ops
    with pytest.raises(_native.NativeError):
        ops.grid_subsample(torch.from_numpy(p), torch.from_numpy(l), 0.3)       # CPU tensors are refused
