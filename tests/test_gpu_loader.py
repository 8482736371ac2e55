"""GPU tests (-m gpu) of the device CSR assembly behind the file loaders:
sddmm_coo_to_csr against the reference loader's rules (src/Matrix.cpp:398-480: duplicate => reject, stable sort by
ROW ONLY so that file order survives inside a row, row offsets), directly and through `BSMR-sddmm -x 1` with the
loader forced onto the device (SDDMM_B200_LOADER=device) next to the host stages and the reference's own loader."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from cases import ROOT, gen, pkg
from oracle import oracle as O
from test_oracle_cpu import _MTX_QUIRKS, _fnv

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "sddmm-gpu_b200", "BSMR-sddmm")


def _coo_to_csr(rows, cols, vals, M, N):
    L = pkg.lib()
    n = len(rows)
    ro = np.zeros(M + 1, np.uint32)
    ci = np.zeros(max(1, n), np.uint32)
    va = np.zeros(max(1, n), np.float32)
    dup = C.c_int(0)
    rc = L.sddmm_coo_to_csr(rows.ctypes.data, cols.ctypes.data, vals.ctypes.data if vals is not None else None, n, M, N,
                            ro.ctypes.data, ci.ctypes.data, va.ctypes.data if vals is not None else None, C.byref(dup))
    assert rc == 0, L.sddmm_last_error()
    return ro, ci[:n], va[:n], bool(dup.value)


@pytest.mark.parametrize("shape", [(1, 7, 5), (64, 80, 400), (1000, 33, 5000), (5000, 70000, 300000)])
def test_device_csr_build_is_the_stable_row_sort(shape):
    M, N, n = shape
    rng = np.random.default_rng(M + N)
    key = rng.choice(M * N, size=min(n, M * N), replace=False)  # unique (row, col), random FILE order
    rows, cols = (key // N).astype(np.uint32), (key % N).astype(np.uint32)
    vals = rng.random(len(key)).astype(np.float32)
    ro, ci, va, dup = _coo_to_csr(rows, cols, vals, M, N)
    assert not dup
    order = np.argsort(rows, kind="stable")  # the reference: thrust::host stable sort by row only (Matrix.cpp:467)
    assert np.array_equal(ci, cols[order]) and np.array_equal(va, vals[order])
    want = np.zeros(M + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=M), out=want[1:])
    assert np.array_equal(ro, want.astype(np.uint32))
    # pattern-only form
    ro2, ci2, _, dup2 = _coo_to_csr(rows, cols, None, M, N)
    assert not dup2 and np.array_equal(ro2, ro) and np.array_equal(ci2, ci)


def test_device_csr_build_flags_duplicates():
    rows = np.array([3, 0, 3, 1, 3], np.uint32)
    cols = np.array([5, 1, 2, 1, 5], np.uint32)
    vals = np.ones(5, np.float32)
    assert _coo_to_csr(rows, cols, vals, 4, 8)[3]
    cols[4] = 6
    assert not _coo_to_csr(rows, cols, vals, 4, 8)[3]


def _loader_line(path, mode):
    env = dict(os.environ, SDDMM_B200_LOADER=mode)
    r = subprocess.run([EXE, "-f", path, "-x", "1"], capture_output=True, text=True, timeout=120, env=env)
    line = [l for l in r.stdout.splitlines() if l.startswith("[loader")]
    return r.returncode, (line[0] if line else None), r.stderr


@pytest.mark.parametrize("name", sorted(_MTX_QUIRKS))
def test_cli_loader_on_device_matches_host_and_reference(tmp_path, name):
    """the 14 accept / reject quirk cases with the CSR assembled on the GPU"""
    assert os.access(EXE, os.X_OK), "CLI not built"
    p = str(tmp_path / (name + ".mtx"))
    with open(p, "w", newline="") as f:
        f.write(_MTX_QUIRKS[name])
    rc_d, line_d, err_d = _loader_line(p, "device")
    rc_h, line_h, _ = _loader_line(p, "host")
    assert (rc_d == 0) == (rc_h == 0), err_d
    assert line_d == line_h
    if O.ref_available():
        rc, res = O.ref_load_mtx(p)
        assert (rc == 0) == (rc_d == 0)
        if rc == 0:
            M, N, ro, ci, va = res
            tok = line_d.strip("[]").split()
            got = dict(zip(tok[2::2], tok[3::2]))
            assert int(got["rowOff"], 16) == _fnv(ro) and int(got["colIdx"], 16) == _fnv(ci) and int(got["values"], 16) == _fnv(va)


def test_cli_loader_on_device_keeps_file_order_inside_rows(tmp_path):
    S = gen.shuffle_within_rows(gen.with_empty_rows(gen.rmat(12, 8, 3), 5), 4)
    p = str(tmp_path / "g.mtx")
    gen.write_mtx(p, S, order="rowrev")
    rc_d, line_d, err = _loader_line(p, "device")
    rc_h, line_h, _ = _loader_line(p, "host")
    assert rc_d == 0 and rc_h == 0 and line_d == line_h, err
