"""The bit-exact parallel fp32 fold (graphmat_b200/include/GraphMat/gm_fadd32.cuh) against a plain
serial fp32 sum: host walk of the algorithm (CPU) and the two device kernels (gpu)."""
import numpy as np
import pytest

from graphmat_b200 import capi


def serial(a):
    a = np.asarray(a, np.float32)
    return np.cumsum(a, dtype=np.float32)[-1]  # numpy's cumsum is a serial left fold


def cases(seed, count, nmax):
    rng = np.random.default_rng(seed)
    for t in range(count):
        n = int(rng.integers(1, nmax))
        kind = t % 7
        if kind == 0:
            a = rng.random(n) * 1e-6
        elif kind == 1:  # exact ties against the accumulator's ulp
            a = rng.integers(1, 8, n) * 2.0 ** -20
        elif kind == 2:  # huge dynamic range: many binade crossings
            a = np.exp(rng.normal(0, 4, n))
        elif kind == 3:  # zeros sprinkled in (PageRank sends 0 for degree-0 vertices)
            a = rng.random(n)
            a[rng.integers(0, n, max(1, n // 50))] = 0
        elif kind == 4:
            a = (rng.random(n) < 0.5) * 0.5 + 2.0 ** -24
        elif kind == 5:  # negatives: outside the precondition, must still be exact through the serial path
            a = rng.normal(0, 1, n)
        else:  # PageRank-like messages
            a = 0.3 / rng.integers(1, 100000, n)
        yield np.asarray(a, np.float32)


def test_numpy_cumsum_is_serial():
    a = np.random.default_rng(0).random(3000).astype(np.float32)
    s = np.float32(0)
    for v in a:
        s = np.float32(s + v)
    assert s == serial(a)


def test_host_fold_is_bit_exact():
    for a in cases(0, 210, 6000):
        assert capi.fold_f32_host(a).tobytes() == serial(a).tobytes()
    for a in cases(1, 14, 300000):
        assert capi.fold_f32_host(a).tobytes() == serial(a).tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("warps", [1, 16, 32])  # 32: the TMA-streamed kernel of the staged longest rows
def test_device_fold_is_bit_exact(warps):
    for i, a in enumerate(cases(2, 70, 20000)):
        r = capi.fold_f32_device(a, warps=warps, offset=i % 9)
        e = serial(a)
        assert r.tobytes() == e.tobytes() or (np.isnan(r) and np.isnan(e)), (i, len(a), r, e)
    for i, a in enumerate(cases(3, 14, 1500000)):
        r = capi.fold_f32_device(a, warps=warps, offset=(3 * i) % 8)
        assert r.tobytes() == serial(a).tobytes(), (i, len(a))
