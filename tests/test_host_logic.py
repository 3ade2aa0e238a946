"""CPU tests of the host side: wavelet descriptors (WT mirror), Util helpers, layout helpers, the C-ABI
library's exported surface and its argument checking (no GPU work is launched)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import wavelets_b200 as wb
from wavelets_b200 import WT, wavelet, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ortho_filters_are_orthonormal():
    for c in [WT.haar, WT.db2, WT.db4, WT.db10, WT.coif4, WT.sym8, WT.beyl, WT.vaid]:
        h = wavelet(c).qmf
        assert abs(np.dot(h, h) - 1) < 1e-12
        for k in range(1, len(h) // 2):                      # double-shift orthogonality
            tol = 2e-8 if c.namebase in ("coif", "sym", "beyl", "vaid") else 1e-10   # tabulated to ~10 digits
            assert abs(np.dot(h[2 * k:], h[:-2 * k])) < tol, (c, k)
    assert abs(wavelet(WT.db4).qmf.sum() - np.sqrt(2)) < 1e-10


def test_daubechies_db2_closed_form():
    s3 = np.sqrt(3.0)
    ref = np.array([1 + s3, 3 + s3, 3 - s3, 1 - s3]) / (4 * np.sqrt(2))
    assert np.max(np.abs(wavelet(WT.db2).qmf - ref)) < 1e-14
    assert len(wavelet(WT.db6)) == 12 and len(wavelet(WT.batt6)) == 59


def test_wavelet_constructor_and_errors():
    f = wavelet(WT.db4)
    assert f.name == "db4" and len(f) == 8
    g = wavelet(WT.cdf97, WT.Lifting)
    assert g.name == "cdf9/7" and len(g.step) == 4 and g.step[0].steptype == "update"
    assert wavelet(WT.db2, WT.Lifting).step[2].shift == -1
    with pytest.raises(TypeError):
        wavelet(WT.cdf97)                      # no filter form
    with pytest.raises(ValueError):
        wavelet(WT.db4, WT.Lifting)            # "scheme not found"
    with pytest.raises(ValueError):
        wavelet(WT.Coiflet(3))                 # "filter not found"


def test_util_helpers():
    assert wb.maxtransformlevels(1024) == 10 and wb.maxtransformlevels(24) == 3 and wb.maxtransformlevels(1) == 0
    assert wb.maxtransformlevels(torch.empty(16, 24)) == 3
    assert wb.detailindex(1024, 1, 1) == 513 and wb.detailn(1024, 3) == 128
    assert list(wb.detailrange(16, 2)) == [5, 6, 7, 8]
    t = wb.maketree(16, 2, "full")
    assert t.tolist() == [1, 1, 1] + [0] * 12 and wb.isvalidtree(16, t)
    assert wb.maketree(16, 3, "dwt").tolist() == [1, 1, 0, 1] + [0] * 11
    bad = np.zeros(15, dtype=np.uint8); bad[1] = 1
    assert not wb.isvalidtree(16, bad) and not wb.isvalidtree(16, np.ones(7))
    # odd length: maxtransformlevels = 0, the only valid tree is the empty one (no TypeError from a fractional range)
    assert wb.isvalidtree(7, np.zeros(0, dtype=np.uint8)) and not wb.isvalidtree(7, np.zeros(1, dtype=np.uint8))


def test_colmajor_layout_helper():
    a = torch.arange(24.0).reshape(4, 6)
    c = wb.colmajor(a)
    assert c.stride() == (1, 4) and torch.equal(c, a)
    assert wb.colmajor(c).data_ptr() == c.data_ptr()
    v = torch.arange(5.0)
    assert wb.colmajor(v).data_ptr() == v.data_ptr()
    b = torch.arange(24.0).reshape(2, 3, 4)
    assert wb.colmajor(b).stride() == (1, 2, 6)


def test_no_cpu_fallback():
    with pytest.raises(TypeError, match="no CPU path"):
        wb.dwt(torch.randn(16), wavelet(WT.db2))


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "wavelets_b200.h")).read()
    declared = set(re.findall(r"\b(wb200_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = _lib.lib()
    for s in declared:
        assert hasattr(L, s), s
    assert L.wb200_version() == 100


def test_abi_argument_checks_without_gpu():
    """The reference's validation order and messages (transforms_filter.jl:25-34) come back as status codes
    before any CUDA call is made."""
    L = _lib.lib()
    q = np.ascontiguousarray(wavelet(WT.db2).qmf)
    qp = q.ctypes.data_as(C.POINTER(C.c_double))
    fake_x, fake_y = 0x1000, 0x2000
    call = lambda y, x, n, Lv: L.wb200_dwt_filter(y, x, 1, _lib.dims_array([n]), 1, qp, 4, Lv, 1, _lib.F64, None, 0, None, 0)
    assert call(fake_y, fake_x, 24, -1) == _lib.ELEVEL
    assert L.wb200_status_string(_lib.ELEVEL) == b"L must be positive"
    assert call(fake_y, fake_x, 24, 4) == _lib.EPOW2
    assert call(fake_x, fake_x, 24, 1) == _lib.EALIAS
    assert b"in array is out array" in L.wb200_last_error_string()
    assert L.wb200_dwt_filter(fake_y, fake_x, 4, _lib.dims_array([8]), 1, qp, 4, 1, 1, _lib.F64, None, 0, None, 0) == _lib.EDIMS
    assert L.wb200_dwt_filter(fake_y, fake_x, 1, _lib.dims_array([8]), 1, qp, 4, 1, 1, 9, None, 0, None, 0) == _lib.EDTYPE
    assert L.wb200_dwt_filter(fake_y, fake_x, 1, _lib.dims_array([8]), 1, qp, 1, 1, 1, _lib.F64, None, 0, None, 0) == _lib.EARG
    steps, ns = _lib.make_steps(wavelet(WT.db2, WT.Lifting))
    assert L.wb200_dwt_lifting(fake_y, fake_x, 2, _lib.dims_array([8, 16]), 1, steps, ns, 1.0, 1.0, 1, 1, _lib.F32,
                               None, 0, None, 0) == _lib.ENOTCUBE
    tree = np.array([0, 1, 0], dtype=np.uint8)
    assert L.wb200_wpt_filter(fake_y, fake_x, 4, 1, qp, 4, tree.ctypes.data_as(C.POINTER(C.c_uint8)), 3, 1, _lib.F64,
                              None, 0, None, 0) == _lib.ETREE
    assert L.wb200_maxtransformlevels(1 << 20) == 20
    assert L.wb200_workspace_bytes(0, 1, _lib.dims_array([1024]), 4, 10, _lib.F32, _lib.FLAG_FORCE_GENERIC) >= (512 + 256) * 4 * 4
    assert L.wb200_launch_count(1) == 0


def test_shard_ranges():
    from wavelets_b200.shard import shard_range, shard_sizes
    assert shard_sizes(10, 4) == [3, 3, 2, 2]
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_range(1024, 7, 8) == (896, 1024)
    assert sum(shard_sizes(5, 8)) == 5


def test_widened_entry_points_argument_checks_without_gpu():
    """MODWT / threshold / denoise / best-basis entry points validate before any CUDA call (status codes carry the
    reference's messages: transforms_maximal_overlap.jl:46-48, denoising.jl:30, entropy.jl:52)."""
    L = _lib.lib()
    q = np.ascontiguousarray(wavelet(WT.db4).qmf)
    qp = q.ctypes.data_as(C.POINTER(C.c_double))
    fx, fy = 0x1000, 0x2000
    assert L.wb200_modwt(fy, fx, 129, 1, qp, 8, 8, _lib.F64, None, 0, None, 0) == _lib.ELEVEL
    assert b"Too many transform levels" in L.wb200_last_error_string()
    assert L.wb200_modwt(fy, fx, 129, 1, qp, 8, 0, _lib.F64, None, 0, None, 0) == _lib.ELEVEL
    assert b"L must be >= 1" in L.wb200_last_error_string()
    assert L.wb200_modwt(fy, fx, 129, 1, qp, 8, 3, _lib.C64, None, 0, None, 0) == _lib.EDTYPE
    assert L.wb200_maxmodwttransformlevels(129) == 7 and L.wb200_maxmodwttransformlevels(128) == 7
    assert L.wb200_threshold(fx, 10, 9, 1.0, _lib.F32, None) == _lib.EARG
    assert L.wb200_threshold(fx, 10, 0, -1.0, _lib.F32, None) == _lib.EARG                      # @assert t >= 0
    assert L.wb200_threshold(fx, 0, 0, 1.0, _lib.F32, None) == _lib.OK                          # empty array: nothing to do
    assert L.wb200_threshold_biggest(fx, 10, -1, _lib.F32, None) == _lib.EARG                   # @assert m >= 0
    assert L.wb200_threshold_biggest(fx, 10, 10, _lib.F32, None) == _lib.OK                     # m >= n keeps everything
    spin = (C.c_int32 * 3)(8, 8, 1)
    null_s = C.POINTER(_lib.LiftStep)()
    nan = float("nan")
    den = lambda dims, nd, wk, TI: L.wb200_denoise(fy, fx, nd, _lib.dims_array(dims), wk, qp, 8, null_s, 0, 0.0, 0.0, 2, 0, 3.0, nan,
                                                   TI, spin, _lib.F32, None, 0)
    assert den([16, 32], 2, 1, 0) == _lib.ENOTCUBE and b"square/cube" in L.wb200_last_error_string()
    assert den([16], 1, 0, 1) == _lib.EARG and b"TI not supported" in L.wb200_last_error_string()
    assert den([16], 1, 7, 0) == _lib.EARG
    tree = np.array([0, 1, 0], dtype=np.uint8)
    best = np.zeros(3, dtype=np.uint8)
    pu8 = C.POINTER(C.c_uint8)
    assert L.wb200_bestbasistree(best.ctypes.data_as(pu8), None, None, fx, 4, 1, qp, 8, null_s, 0, 0.0, 0.0, tree.ctypes.data_as(pu8), 3, 0,
                                 _lib.F64, None, 0) == _lib.ETREE
    out = C.c_double(1.0)
    assert L.wb200_coefentropy(C.byref(out), fx, 0, 0, nan, _lib.F64, None) == _lib.OK and out.value == 0.0
    assert L.wb200_coefentropy(C.byref(out), fx, 4, 5, nan, _lib.F64, None) == _lib.EARG


def test_threshold_module_host_side():
    assert abs(wb.VisuShrink(256).t - np.sqrt(2 * np.log(256))) < 1e-15 and isinstance(wb.VisuShrink(256).th, wb.HardTH)
    assert wb.VisuShrink(wb.SoftTH(), 2.5).t == 2.5
    assert len(wb.Threshold.DEFAULT_WAVELET.qmf) == 10                                           # sym5
    assert [t.kind for t in (wb.HardTH(), wb.SoftTH(), wb.SemiSoftTH(), wb.SteinTH(), wb.NegTH(), wb.PosTH())] == [0, 1, 2, 3, 4, 5]
    assert (wb.ShannonEntropy().kind, wb.LogEnergyEntropy().kind) == (0, 1)
    for fn in (lambda: wb.threshold(torch.randn(8), wb.HardTH(), 1.0), lambda: wb.denoise(torch.randn(8)),
               lambda: wb.noisest(torch.randn(8)), lambda: wb.coefentropy(torch.randn(8)),
               lambda: wb.bestbasistree(torch.randn(8), wavelet(WT.db2)), lambda: wb.modwt(torch.randn(8), wavelet(WT.db2))):
        with pytest.raises(TypeError, match="no CPU path"):
            fn()


def test_fused_packet_level_index_arithmetic():
    """tools/wptfused_emul.py: the plans, band offsets, rotated stores and staged ranges of the fused K-level packet kernels
    (wptfused.cu) emulated in numpy against a direct periodic packet transform -- a CPU dry run of the index arithmetic the GPU
    tests then check bit for bit"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("wptfused_emul", os.path.join(ROOT, "tools", "wptfused_emul.py"))
    em = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(em)
    for F, K, tile, nj, PA, vec in [(16, 4, 256, 1024, 2, 4), (8, 3, 256, 1024, 2, 4), (18, 4, 256, 1024, 1, 2), (12, 2, 256, 768, 2, 4)]:
        h, g = em.qmf_pair(F)
        x = em.rng.standard_normal(nj)
        ref = em.wpt_ref(x, h, g, K)
        assert np.max(np.abs(ref - em.emul_ana(x, h, g, K, tile, PA, vec))) < 1e-9
        cur = np.split(ref, 1 << K)
        for _ in range(K):
            cur = [em.syn_ref(cur[2 * i], cur[2 * i + 1], h, g) for i in range(len(cur) // 2)]
        assert np.max(np.abs(cur[0] - em.emul_syn(ref, h, g, K, tile))) < 1e-9
