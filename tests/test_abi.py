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
