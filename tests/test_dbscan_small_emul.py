"""CPU: the phase code of the one-launch small-input DBSCAN1D kernel (contextsv_b200/csrc/dbscan_small.h), compiled for
the host and run phase by phase, against the golden vectors of the compiled reference and against the oracle on random
inputs.  (The device runs the same phases with one thread per point; tests/test_gpu_parity.py covers that side.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import util
from contextsv_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("dbs") / "libdbs.so")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I" + build.CSRC, os.path.join(ROOT, "tests", "native", "dbscan_small_emul.cpp"), "-o", so]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout[-3000:]
    L = C.CDLL(so)
    for f in (L.dbs_emul, L.dbs_emul_reversed):
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    return L


def fit(f, pts, eps, mp):
    pts = np.ascontiguousarray(pts, np.int32)
    lab = np.full(len(pts), -77, np.int32); nc = np.zeros(1, np.int32)
    assert f(pts.ctypes.data, len(pts), float(eps), int(mp), lab.ctypes.data, nc.ctypes.data) == 0
    return lab, int(nc[0])


def test_golden_vectors(emul):
    n = 0
    for i, pts, eps, mp, labels, largest in util.golden_db_cases():
        if len(pts) == 0 or len(pts) > 1024 or not (eps >= 0):
            continue
        for f in (emul.dbs_emul, emul.dbs_emul_reversed):
            lab, nc = fit(f, pts, eps, mp)
            assert np.array_equal(lab, labels), "case %d eps=%g minPts=%d" % (i, eps, mp)
            assert nc == (int(labels.max()) + 1 if (labels >= 0).any() else 0)
        n += 1
    assert n >= 80


def test_random_against_the_oracle(emul, oracle):
    rng = np.random.default_rng(2026)
    for trial in range(400):
        n = int(rng.choice([1, 2, 3, 5, 17, 64, 200, 1024]))
        span = int(rng.choice([5, 50, 1000, 100000]))
        pts = rng.integers(-span, span, n).astype(np.int32)
        if trial % 7 == 0:
            pts[: n // 2] = pts[0]                                   # pile of duplicates
        if trial % 11 == 0:
            pts = np.concatenate([pts[: n // 2], [np.iinfo(np.int32).max, np.iinfo(np.int32).min]]).astype(np.int32)[:1024]
        eps = float(rng.choice([0, 0.5, 1, 2.9, 10, 100, 1e5, 5e9]))
        mp = int(rng.choice([-1, 0, 1, 2, 3, 5, 10, 2000]))
        want = oracle.dbscan1d(pts, eps, mp)
        for f in (emul.dbs_emul, emul.dbs_emul_reversed):
            lab, _ = fit(f, pts, eps, mp)
            assert np.array_equal(lab, want), "trial %d n=%d eps=%g minPts=%d" % (trial, len(pts), eps, mp)
