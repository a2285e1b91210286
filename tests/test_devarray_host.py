# -*- coding: utf-8 -*-
"""Host-side logic of the `cupy` shim's array type (indexing / assignment semantics the
reference's drivers rely on), exercised on CPU tensors.  The engine itself needs CUDA."""
import numpy as np
import pytest
import torch

from adi_thermal_fields_b200 import devarray as cp


@pytest.fixture(autouse=True)
def cpu_arrays():
    cp._FORCE_DEVICE = torch.device("cpu")
    yield
    cp._FORCE_DEVICE = None


def test_creation_and_transfer():
    T = cp.full((3, 4, 5), 20.0, dtype=cp.float64)      # quick_compare_neumann_robin_backend.py:140
    assert T.shape == (3, 4, 5) and T.dtype == np.float64
    assert np.array_equal(cp.asnumpy(T), np.full((3, 4, 5), 20.0))
    m = cp.asarray(np.ones((3, 4, 5), bool))
    assert m.dtype == np.bool_ and m.get().all()
    z = cp.zeros_like(m, dtype=cp.bool_)
    assert not z.get().any()


def test_birth_mutations():
    # quick_compare_layer_birth_robin_v3.py:272-277
    mask = cp.asarray(np.zeros((4, 4, 6), bool))
    T = cp.full((4, 4, 6), 20.0, dtype=cp.float64)
    born = cp.zeros_like(mask, dtype=cp.bool_)
    cross = cp.asarray(np.eye(4, dtype=bool), dtype=cp.bool_)
    for kk in range(2, 4):
        born[:, :, kk] = cross
    v0 = mask._t._version
    T[born] = 1000.0
    mask[born] = True
    assert mask._t._version > v0           # the engine keys the neighbour code on this
    assert int(mask.sum()) == 8 and float(T.max()) == 1000.0
    assert np.array_equal(mask.get()[:, :, 2], np.eye(4, dtype=bool))
    # waam_from_stl_v7_mm.py:491-493: index tuple from np.where
    idx = np.where(np.eye(4, dtype=bool)[:, :, None] & np.ones((4, 4, 6), bool))
    T[idx] = np.float64(5.0)
    assert float(T.min()) == 5.0


def test_slicing_and_probe_reads():
    a = np.arange(60, dtype=float).reshape(3, 4, 5)
    T = cp.asarray(a)
    assert np.array_equal(cp.asnumpy(T[1, 2, :]), a[1, 2, :])
    assert np.array_equal(cp.asnumpy(T[1:, 2, :]), a[1:, 2, :])
    assert float(cp.asnumpy(T[1, 2, 3])) == a[1, 2, 3]
    T[...] = cp.asarray(a * 2)
    assert np.array_equal(T.get(), a * 2)


def test_where_and_logic():
    a = cp.asarray(np.array([1.0, 2.0, 3.0]))
    m = cp.asarray(np.array([True, False, True]))
    assert np.array_equal(cp.where(m, a, 0.0).get(), [1.0, 0.0, 3.0])
    assert np.array_equal((~m).get(), [False, True, False])
    assert np.array_equal((m & ~m).get(), [False] * 3)
    cp.cuda.Stream.null  # attribute exists


def test_numpy_type_promotion():
    """Python float x bool / int array -> float64 (torch alone gives float32), true division of
    integer arrays -> float64: the reference's CuPy code relies on it (adi3d_gpu_coeff.py:175-176)."""
    import torch
    from adi_thermal_fields_b200 import devarray as cp
    old = cp._FORCE_DEVICE
    cp._FORCE_DEVICE = torch.device("cpu")
    try:
        m = cp.asarray(np.array([True, False, True]))
        i = cp.asarray(np.array([1, 2, 3]))
        x = 0.1 + 1e-12
        for got, want in [((-x) * m, (-x) * np.array([True, False, True])), (m * x, np.array([True, False, True]) * x),
                          (i * x, np.array([1, 2, 3]) * x), (x - i, x - np.array([1, 2, 3])),
                          (i / 3, np.array([1, 2, 3]) / 3), (1.0 / i, 1.0 / np.array([1, 2, 3])),
                          (np.float64(x) * m, np.float64(x) * np.array([True, False, True]))]:
            assert got.dtype == np.float64
            assert np.array_equal(cp.asnumpy(got), want)
    finally:
        cp._FORCE_DEVICE = old
