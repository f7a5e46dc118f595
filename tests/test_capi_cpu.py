"""CPU-side checks of the drop-in boundary: the library loads without a GPU, exports every symbol the header
declares, and every compute entry point refuses to run without an sm_100 device (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

from mimo_b200 import lib as L
from mimo_b200 import ops

ROOT = Path(__file__).resolve().parents[1]


def _header_functions():
    text = (ROOT / "include" / "mimo_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mimo_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    names = _header_functions()
    assert len(names) >= 18
    cdll = ctypes.CDLL(str(L.LIB_PATH))
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in include/mimo_b200.h but not exported by the library"
        assert n in L.SYMBOLS, f"{n} declared in the header but has no ctypes binding in mimo_b200/lib.py"
    for n in L.SYMBOLS:
        assert n in names, f"{n} is bound in lib.py but not declared in the header"


def test_version_and_struct_layout():
    lib = L.load()
    assert b"sm_100a" in lib.mimo_version()
    # the ctypes mirrors must match the C structs (pointer + int64 + float layout, natural alignment)
    for which, st in enumerate((L.Epilogue, L.GemmParams, L.Conv3x3Params, L.GroupNormParams, L.AttnParams)):
        assert lib.mimo_abi_sizeof(which) == ctypes.sizeof(st), st.__name__
    assert lib.mimo_abi_sizeof(99) == -1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    lib = L.load()
    assert lib.mimo_device_check(0) != 0
    assert b"no CPU fallback" in lib.mimo_last_error() or b"CUDA" in lib.mimo_last_error()
    a = torch.zeros(8, 8, dtype=torch.float16)
    with pytest.raises(L.MimoError):
        ops.gemm(a, a)
    from mimo_b200.host.modules import PoseGuider
    pg = PoseGuider(320, block_out_channels=(16, 32, 96, 256)).half()
    with pytest.raises(L.MimoError):
        pg(torch.zeros(1, 3, 1, 64, 64, dtype=torch.float16))


def test_geglu_granule_and_arg_validation():
    lib = L.load()
    assert lib.mimo_gemm_geglu_granule(2560) in (32, 64, 128)
    p = L.GemmParams()
    assert lib.mimo_gemm(ctypes.byref(p), None) == -1  # MIMO_ERR_ARG: null pointers
    assert b"null" in lib.mimo_last_error()


def test_integration_md_stub_binds_the_library():
    """The ctypes stub INTEGRATION.md shows a maintainer (section 2) is executed as written up to its first GPU call: the
    struct mirror must match the C struct, the workspace query must answer on the CPU, and a launch without a device
    must come back as an error string, not a crash."""
    import os
    text = (ROOT / "INTEGRATION.md").read_text()
    block = re.search(r"```python\n(# src/models/resnet\.py.*?)```", text, flags=re.S).group(1)
    ns = {}
    cwd = os.getcwd()
    os.chdir(ROOT)  # the stub loads "mimo_b200/libmimo_b200.so" relative to the checkout
    try:
        exec(compile(block, "INTEGRATION.md:stub", "exec"), ns)
    finally:
        os.chdir(cwd)
    lib, P = ns["_lib"], ns["GroupNormParams"]
    p = P(1 << 20, 320, None, 0, 1 << 20, 1 << 20, 1 << 20, None, 2, 4096, 32, 1e-5, 1, 0)
    need = lib.mimo_groupnorm_workspace_bytes(ctypes.byref(p))
    assert need > 0 and need % 8 == 0
    bad = P(1 << 20, 321, None, 0, 1 << 20, 1 << 20, 1 << 20, None, 2, 4096, 32, 1e-5, 1, 0)
    assert lib.mimo_groupnorm_workspace_bytes(ctypes.byref(bad)) < 0 and b"multiples of 8" in lib.mimo_last_error()
    if not torch.cuda.is_available():
        p.stats = 1 << 20
        assert lib.mimo_groupnorm(ctypes.byref(p), None) < 0
        assert b"CUDA" in lib.mimo_last_error() or b"fallback" in lib.mimo_last_error()
