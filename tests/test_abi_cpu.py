"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol the header
declares, validates arguments like the reference constructors, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def p():
    import pdmpflux_b200
    if not os.path.exists(pdmpflux_b200.LIB_PATH):
        pdmpflux_b200.build()
    return pdmpflux_b200


def test_header_and_binding_agree(p):
    hdr = open(os.path.join(ROOT, "include", "pdmpflux_cuda.h")).read()
    declared = set(re.findall(r"\b(pdmpflux_[a-z_0-9]+)\s*\(", hdr))
    from pdmpflux_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    l = C.CDLL(p.LIB_PATH)
    for name in declared:
        assert hasattr(l, name), name
    assert p.lib().pdmpflux_version() == 200


def test_struct_layouts_match_header(p):
    from pdmpflux_b200 import _lib
    assert C.sizeof(_lib.Config) == 10 * 4 + 4 * 8
    assert C.sizeof(_lib.Tape) == 3 * 8 + 3 * 8 + 8
    assert C.sizeof(_lib.History) == 12 * 8 + 8 + 8 + 8   # is_active appended in 0.2.0


def test_constructor_validation_and_rewrites(p):
    with pytest.raises(p.ArgumentError, match="dimension dim must be positive"):
        p.ZigZag(0, p.GaussStd())
    with pytest.raises(p.ArgumentError, match="grid_size must be non-negative"):
        p.BPS(3, p.GaussStd(), grid_size=-2)
    with pytest.raises(p.ArgumentError):
        p.ZigZag(3, p.GaussStd(), grid_size=1)
    with pytest.raises(p.ArgumentError, match="at least 2"):
        p.ForwardECMC(1, p.GaussStd())
    with pytest.raises(p.ArgumentError, match="Unsupported AD_backend"):
        p.ZigZag(3, p.GaussStd(), AD_backend="Tapenade")
    with pytest.raises(p.UnsupportedError):
        p.ZigZag(3, lambda x: x)
    with pytest.raises(p.UnsupportedError):
        p.ForwardECMC(3, p.GaussStd(), normal=True)
    with pytest.raises(p.DimensionMismatch):
        p.ZigZag(3, p.GaussDiag([1.0, 2.0]))
    # ZigZagSamplers.jl:73-78: tmax == 0 -> (1.0, adaptive); signed && !vectorized -> unsigned (+ warning)
    with pytest.warns(UserWarning, match="Signed bound"):
        s = p.ZigZag(3, p.GaussStd(), tmax=0, adaptive=False, vectorized_bound=False)
    assert s.tmax == 1.0 and s.adaptive and not s.signed_bound
    # BPS/Boomerang/FECMC force vectorized_bound=false; FECMC forces refresh 0 and mix_p=0 in dim 2
    assert not p.BPS(3, p.GaussStd(), vectorized_bound=True).vectorized_bound
    f = p.ForwardECMC(2, p.GaussStd())
    assert f.mix_p == 0.0 and f.refresh_rate == 0.0
    assert p.BPS(3, p.GaussStd()).tmax == 1.0 and p.BPS(3, p.GaussStd()).refresh_rate == 0.1
    assert p.BPSAD(3, p.GaussStd()).refresh_rate == 0.0 and p.Boomerang(3, p.GaussStd()).flow_kind == 1


def test_driver_validation_without_gpu(p):
    s = p.ZigZag(3, p.GaussStd())
    with pytest.raises(p.ArgumentError, match="n_sk must be positive"):
        p.sample_skeleton(s, 0, np.zeros(3), np.ones(3))
    with pytest.raises(p.DimensionMismatch):
        p.sample_skeleton(s, 5, np.zeros(4), np.ones(4))
    with pytest.raises(p.ArgumentError):
        p.sample_skeleton(p.ZigZag(1, p.GaussStd()), 5, float("nan"), 1.0)


def test_no_cpu_fallback(p):
    n = C.c_int(-1)
    p.lib().pdmpflux_device_count(C.byref(n))
    if n.value > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(p.CudaError, match="no CPU fallback|CUDA"):
        p.sample_skeleton(p.ZigZag(3, p.GaussStd()), 5, np.zeros(3), np.ones(3), seed=1)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pdmpflux.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle_c" not in src and "pdmp_oracle" not in src and "libpdmp_oracle" not in src, f


def test_reduction_entry_points_without_gpu(p):
    """pdmpflux_moments_reduce / pdmpflux_comm_* / pdmpflux_moments_allreduce (kernel K5): argument
    validation, the single-rank communicator (needs neither NCCL nor a GPU) and the no-fallback rule."""
    from pdmpflux_b200 import _lib
    lib = p.lib()
    h = C.c_void_p()
    with pytest.raises(p.ArgumentError):
        _lib.check(lib.pdmpflux_comm_create(None, 0, 0, C.byref(h)))
    with pytest.raises(p.ArgumentError):
        _lib.check(lib.pdmpflux_comm_create(None, 2, 2, C.byref(h)))
    with pytest.raises(p.ArgumentError):
        _lib.check(lib.pdmpflux_comm_create(None, 2, 0, C.byref(h)))       # more than one rank needs the unique id
    _lib.check(lib.pdmpflux_comm_create(None, 1, 0, C.byref(h)))            # one rank: no NCCL involved
    buf = np.arange(8, dtype=np.float64)
    _lib.check(lib.pdmpflux_moments_allreduce(h, buf.ctypes.data, 8, None))  # no-op on a single rank
    assert np.array_equal(buf, np.arange(8.0))
    with pytest.raises(p.ArgumentError):
        _lib.check(lib.pdmpflux_moments_allreduce(h, None, 8, None))
    _lib.check(lib.pdmpflux_comm_destroy(h))
    with pytest.raises(p.ArgumentError):
        _lib.check(lib.pdmpflux_comm_unique_id(None, 128))
    m = np.ones((4, 3)); out = np.zeros((4, 3))
    with pytest.raises(p.ArgumentError):
        _lib.check(lib.pdmpflux_moments_reduce(0, 4, m.ctypes.data, m.ctypes.data, None, out.ctypes.data, 0, None))
    n = C.c_int(-1)
    lib.pdmpflux_device_count(C.byref(n))
    if n.value <= 0:
        with pytest.raises(p.CudaError, match="no CPU fallback|CUDA"):
            _lib.check(lib.pdmpflux_moments_reduce(3, 4, m.ctypes.data, m.ctypes.data, None, out.ctypes.data, 0, None))
    # the python wrapper builds a single-rank communicator when torch.distributed is not initialised
    c = p.dist.Comm()
    assert (c.rank, c.world) == (0, 1)
    c.close()
