"""CPU: the C-ABI libraries load and export every symbol include/contextsv_b200.h declares
(no compute calls: there is no GPU here), and the product refuses to run without one."""
import ctypes as C
import os
import re

import pytest

from contextsv_b200 import _capi, build


def declared_functions():
    hdr = open(os.path.join(build.INC, "contextsv_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(csv_[a-z0-9_]+)\s*\(", hdr))
    return names


def test_header_and_binding_lists_agree():
    assert declared_functions() == set(_capi.EXPORTS) | set(_capi.SYNTH_EXPORTS)


def test_cuda_library_exports_every_symbol():
    assert os.path.exists(build.LIB_CUDA), "libcontextsv_b200.so not built: python -m contextsv_b200.build"
    L = C.CDLL(build.LIB_CUDA)
    for name in _capi.EXPORTS:
        assert hasattr(L, name), name


def test_synth_library_exports_every_symbol():
    build.build_synth()
    L = C.CDLL(build.LIB_SYNTH)
    for name in _capi.SYNTH_EXPORTS:
        assert hasattr(L, name), name


def test_sm100a_code_is_embedded():
    """The library carries sm_100a SASS and nothing else (no multi-arch fallback)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", build.LIB_CUDA], stdout=subprocess.PIPE, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from contextsv_b200.api import Context
    with pytest.raises(_capi.CsvError) as e:
        Context(0)
    assert e.value.status == 1 and "no CPU fallback" in str(e.value)


def test_host_widen_helper():
    """csv_host_widen_u8 (host half of the narrow depth fetch): every length / alignment, nothing written outside."""
    import numpy as np
    L = _capi.lib()
    rng = np.random.default_rng(3)
    for n in (0, 1, 7, 8, 31, 32, 33, 100, 4097, 1 << 18):
        for off in range(9):
            for soff in (0, 1, 3):
                src = rng.integers(0, 256, n + soff + 16, dtype=np.uint8)[soff:soff + n]
                buf = np.full(n + off + 16, 0xdeadbeef, np.uint32)
                dst = buf[off:off + n]
                L.csv_host_widen_u8(src.ctypes.data, dst.ctypes.data, n)
                assert np.array_equal(dst, src), (n, off, soff)
                assert (buf[:off] == 0xdeadbeef).all() and (buf[off + n:] == 0xdeadbeef).all()
