"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/lv_capi.h declares (no compute calls without a GPU), and the host mirror fails loudly
when the device is missing instead of falling back to anything."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "lv_capi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lv_[A-Za-z_0-9]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol(lv):
    path = lv.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 25
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    from lvb200 import _capi
    assert sorted(_capi.SYMBOLS) == declared  # the ctypes binding covers the whole header


def test_sass_is_sm100a_only(lv):
    """The shipped library carries sm_100a code and nothing else (no multi-arch fallbacks)."""
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", lv.library_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu(lv):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lv.LvError) as ei:
        lv.VoronoiGrid(lv.Rectangle((0, 0), (1, 1)), 0.1)
    assert ei.value.status == 4  # LV_ECUDA


def test_argument_errors_match_reference(lv):
    with pytest.raises(ValueError, match="h must be positive"):  # neighborlist.jl:19-21
        lv.VoronoiGrid(lv.Rectangle((0, 0), (1, 1)), 0.1, h=-1.0)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "lagrangianvoronoi.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().lower(), os.path.join(dirpath, f)


def test_synthetic_generator_is_deterministic(lv):
    import numpy as np
    a = lv.synthetic.jittered_lattice(16, 0)
    b = lv.synthetic.jittered_lattice(16, 0)
    c = lv.synthetic.jittered_lattice(16, 1)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a.shape == (256, 2) and a.min() > 0 and a.max() < 1
    # splitmix64 known answers (mix64(1), mix64(2) from the published finaliser)
    z = lv.synthetic.mix64(np.array([0, 1], dtype=np.uint64))
    assert int(z[0]) == 0 and int(z[1]) == 0x5692161D100B05E5
    # strip extraction keeps the global order
    s = lv.synthetic.jittered_lattice(16, 0, rows=(4, 8))
    full = a.reshape(16, 16, 2)[:, 4:8].reshape(-1, 2)
    assert np.array_equal(s, full)


def test_wire_format_decoder_rebuilds_the_oracles_edge_records(lv, oracle):
    """lv_wire_expand (host-only part of the pipelined download): the 20 B/edge wire format -- start vertex + label word
    with wall / end-of-row bits -- decodes to exactly the 40-byte Edge records, for any split of the edge list into chunks."""
    import numpy as np
    from lvb200._capi import EDGE_DTYPE, load_library, ptr
    from .conftest import make_points
    L = load_library()
    for kind, n_side, per in (("jitter", 40, True), ("poisson", 24, False)):
        xy, dr, bmin, bmax = make_points(kind, n_side, 1)
        og = oracle.OracleGrid(bmin, bmax, dr, xperiodic=per, yperiodic=per)
        og.set_points(xy)
        assert og.remesh() == 0
        rowptr, edges = og.mesh()
        nnz = int(rowptr[-1])
        v1 = np.ascontiguousarray(edges["v1"])
        lab = edges["label"]
        word = np.where(lab > 0, lab, (1 << 30) | (-lab)).astype(np.uint32)
        word[rowptr[1:][np.diff(rowptr) > 0] - 1] |= np.uint32(1 << 31)
        row_of = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
        for chunk in (nnz, 1000, 7):
            out = np.zeros(nnz, EDGE_DTYPE)
            for k0 in range(0, nnz, chunk):
                ln = min(chunk, nnz - k0)
                has_next = k0 + ln < nnz
                vv = np.ascontiguousarray(v1[k0:k0 + ln + has_next])
                ww = np.ascontiguousarray(word[k0:k0 + ln])
                rs = np.ascontiguousarray(v1[rowptr[row_of[k0]]])
                seg = out[k0:k0 + ln]
                assert L.lv_wire_expand(ptr(vv), ptr(ww), ln, int(has_next), ptr(rs), ptr(seg)) == 0
            assert out.tobytes() == edges.tobytes(), (kind, chunk)
    # a list that stops inside a row is refused
    bad = np.zeros(1, np.uint32)
    assert L.lv_wire_expand(ptr(np.zeros(2)), ptr(bad), 1, 0, ptr(np.zeros(2)), ptr(np.zeros(1, EDGE_DTYPE))) != 0


def test_wire_format_decoder_property(lv):
    """Property test of lv_wire_expand: random row structures (empty rows, rows of 3..12 edges, wall labels), random chunk
    boundaries -- the decoded records always equal the records the wire format was derived from."""
    import numpy as np
    from hypothesis import given, settings, strategies as st
    from lvb200._capi import EDGE_DTYPE, load_library, ptr
    L = load_library()

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.sampled_from([0, 0, 3, 4, 5, 6, 7, 9, 12]), min_size=1, max_size=40), st.integers(1, 50), st.integers(0, 2**31 - 1))
    def check(degs, chunk, seed):
        rng = np.random.default_rng(seed)
        degs = np.array(degs)
        rowptr = np.concatenate([[0], np.cumsum(degs)])
        nnz = int(rowptr[-1])
        if nnz == 0:
            return
        edges = np.zeros(nnz, EDGE_DTYPE)
        edges["v1"] = rng.standard_normal((nnz, 2))
        lab = rng.integers(1, 2**30 - 1, nnz)
        wall = rng.random(nnz) < 0.2
        lab[wall] = -rng.integers(1, 5, int(wall.sum()))
        edges["label"] = lab
        for i in range(len(degs)):                                   # closed chains: v2 = successor's v1, last -> first
            a, b = rowptr[i], rowptr[i + 1]
            if b > a:
                edges["v2"][a:b] = np.roll(edges["v1"][a:b], -1, axis=0)
        v1 = np.ascontiguousarray(edges["v1"])
        word = np.where(lab > 0, lab, (1 << 30) | (-lab)).astype(np.uint32)
        word[rowptr[1:][degs > 0] - 1] |= np.uint32(1 << 31)
        row_of = np.repeat(np.arange(len(degs)), degs)
        out = np.zeros(nnz, EDGE_DTYPE)
        for k0 in range(0, nnz, chunk):
            ln = min(chunk, nnz - k0)
            has_next = k0 + ln < nnz
            vv = np.ascontiguousarray(v1[k0:k0 + ln + has_next])
            ww = np.ascontiguousarray(word[k0:k0 + ln])
            rs = np.ascontiguousarray(v1[rowptr[row_of[k0]]])
            seg = out[k0:k0 + ln]
            assert L.lv_wire_expand(ptr(vv), ptr(ww), ln, int(has_next), ptr(rs), ptr(seg)) == 0
        assert out.tobytes() == edges.tobytes()

    check()
