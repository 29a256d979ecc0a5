"""CPU-only checks of the drop-in boundary: libb200itp.so loads, exports every symbol include/b200itp.h
declares, its POD layouts match the ctypes mirrors, and its host-only entry points (system assembly, output
type rule) agree with the oracle / the host mirror.  No kernel is launched here."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200itp.h")


@pytest.fixture(scope="module")
def lib(pkg):
    if not os.path.exists(pkg._lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return pkg._lib.lib()


def test_library_exports_every_declared_symbol(pkg, lib):
    text = open(HEADER).read()
    declared = set(re.findall(r"B200ITP_API\s+[\w\s\*]+?\b(b200itp_\w+)\s*\(", text))
    assert declared == set(pkg._lib.EXPORTED_SYMBOLS)
    out = subprocess.check_output(["nm", "-D", "--defined-only", pkg._lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (b200itp_\w+)", out))
    assert declared <= exported
    assert lib.b200itp_version() == 100


def test_struct_layouts_match_header(pkg):
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "b200itp.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(b200itp_desc), offsetof(b200itp_desc, size), offsetof(b200itp_desc, inner_lb),
         offsetof(b200itp_desc, has_scale), offsetof(b200itp_desc, outer_tag), offsetof(b200itp_desc, etp_kind),
         offsetof(b200itp_desc, fill_value));
  printf("%zu %zu %zu\n", sizeof(b200itp_axis_system), offsetof(b200itp_axis_system, dl), offsetof(b200itp_axis_system, Cp));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        cfile = os.path.join(td, "t.c")
        open(cfile, "w").write(src)
        exe = os.path.join(td, "t")
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), cfile, "-o", exe])
        a, b = subprocess.check_output([exe], text=True).strip().split("\n")
    D, S = pkg._lib.Desc, pkg._lib.AxisSystem
    assert [int(x) for x in a.split()] == [C.sizeof(D), D.size.offset, D.inner_lb.offset, D.has_scale.offset,
                                            D.outer_tag.offset, D.etp_kind.offset, D.fill_value.offset]
    assert [int(x) for x in b.split()] == [C.sizeof(S), S.dl.offset, S.Cp.offset]


@pytest.mark.parametrize("dtype,code", [(np.float32, 0), (np.float64, 1)])
def test_product_system_assembly_matches_oracle(pkg, orc, lib, dtype, code):
    """b200itp_prefiltering_system (host code of the product) against the oracle's restatement of
    prefiltering_system + lut! + Woodbury Cp, for every (degree, bc, grid) with a method in the reference."""
    combos = [(3, 1, 0), (3, 1, 1), (3, 2, 0), (3, 2, 1), (3, 3, 0), (3, 3, 1), (3, 4, 0), (3, 4, 1),
              (2, 1, 0), (2, 1, 1), (2, 2, 0), (2, 2, 1), (2, 3, 0), (2, 3, 1), (2, 4, 0), (2, 4, 1), (2, 5, 0), (2, 5, 1)]
    for n in (7, 40, 515):
        for deg, bc, gt in combos:
            dl = np.zeros(n - 1, dtype)
            d = np.zeros(n, dtype)
            du = np.zeros(n - 1, dtype)
            k = C.c_int32(-1)
            urow = (C.c_int32 * 8)()
            vcol = (C.c_int32 * 8)()
            Cp = np.zeros(64, dtype)
            rc = lib.b200itp_prefiltering_system(code, n, deg, bc, gt, dl.ctypes.data, d.ctypes.data, du.ctypes.data,
                                                 C.byref(k), urow, vcol, Cp.ctypes.data)
            assert rc == 0, pkg._lib.last_error()
            odl, od, odu, our, ovc, oCp = orc.prefiltering_system(dtype, n, deg, bc, gt)
            assert np.array_equal(dl, odl) and np.array_equal(d, od) and np.array_equal(du, odu)
            kk = k.value
            assert kk == len(our)
            pairs = sorted(zip(urow[:kk], vcol[:kk]))
            assert pairs == sorted(zip(our, ovc))
            # same capacitance matrix up to the ordering of the Woodbury terms
            P = Cp[:kk * kk].reshape(kk, kk, order="F")
            perm = [list(zip(urow[:kk], vcol[:kk])).index(p) for p in zip(our, ovc)]
            assert np.allclose(P[np.ix_(perm, perm)], oCp, rtol=1e-4 if dtype == np.float32 else 1e-12, atol=0)
    # no method in the reference: Cubic(Reflect) -> unsupported (axis left unfiltered, prefiltering.jl:124)
    rc = lib.b200itp_prefiltering_system(1, 10, 3, 5, 0, d.ctypes.data, d.ctypes.data, d.ctypes.data, C.byref(k), urow,
                                         vcol, Cp.ctypes.data)
    assert rc == 2


def test_out_dtype_rule_matches_host_mirror(pkg, orc, lib):
    m = pkg
    from interpolations_jl_b200.core import out_dtype_code
    A32 = np.zeros((8, 8), np.float32)
    A64 = np.zeros((8, 8), np.float64)
    cases = []
    for A in (A32, A64):
        for it in (m.BSpline(m.Cubic(m.Line(m.OnGrid()))), m.BSpline(m.Quadratic(m.Reflect(m.OnCell()))), m.BSpline(m.Linear())):
            itp = orc.interpolate(A, it)
            r32 = m.jrange(np.float32(0), np.float32(1), length=8, dtype=np.float32)
            r64 = m.jrange(0.0, 1.0, length=8)
            ri = m.jrange(1, 8)
            for w in (itp, m.scale(itp, r32, r32), m.scale(itp, r64, r64), m.scale(itp, ri, ri)):
                cases += [w, m.extrapolate(w, m.Flat()), m.extrapolate(w, m.Throw()), m.extrapolate(w, 0.0),
                          m.extrapolate(w, 0), m.extrapolate(w, np.float32(0))]
    for w in cases:
        for cdt in (np.float32, np.float64):
            D = w.descriptor(cdt)
            assert lib.b200itp_out_dtype(C.byref(D)) == out_dtype_code(D), (w, cdt)


def test_abi_rejects_bad_arguments_without_a_gpu(pkg, lib):
    D = pkg._lib.Desc()
    D.ndims = 7
    assert lib.b200itp_out_dtype(C.byref(D)) < 0
    k = C.c_int32()
    rc = lib.b200itp_prefiltering_system(5, 10, 3, 1, 0, None, None, None, C.byref(k), None, None, None)
    assert rc > 0 and "null" in pkg._lib.last_error()


def test_product_does_not_import_the_oracle():
    """The product path must not reference oracle/ (no CPU fallback)."""
    pkgdir = os.path.join(ROOT, "interpolations.jl_b200")
    for dirpath, _, files in os.walk(pkgdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".jl")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f == "__init__.py" and False, os.path.join(dirpath, f)
