"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "centernet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cnb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(built_lib)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/centernet_b200.h but not exported: {missing}"


def test_ctypes_table_matches_header(built_lib):
    from centernet_pytorch_lightning_b200 import _lib
    L = _lib.lib()
    sigs = _lib._declare(L)
    assert sorted(sigs) == _declared()
    assert L.cnb_version() == 100


def test_sass_uses_blackwell_paths(built_lib):
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, bulk copy -> UBLKCP (B200_PROFILING.md table)."""
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UBLKCP"):
        assert mnemonic in sass, mnemonic


def test_argument_errors_without_gpu(built_lib):
    """Validation happens before any CUDA call: bad arguments return CNB_ERR_INVALID with a message."""
    from centernet_pytorch_lightning_b200 import _lib
    L = _lib.lib()
    assert L.cnb_ctdet_decode(None, None, None, None, 1, 1, 8, 8, 10, None, 0, None) == -1
    assert b"null" in L.cnb_last_error()
    assert L.cnb_ctdet_decode_workspace_bytes(2, 80, 128, 128, 100) > 0
    assert L.cnb_ctdet_decode_workspace_bytes(2, 80, 128, 128, 100000) == 0
    assert L.cnb_conv_packed_weight_bytes(64, 64, 3, 3) == 64 * 576 * 2


def test_cpu_tensors_fail_loudly(built_lib):
    import pytest
    import torch
    from centernet_pytorch_lightning_b200._lib import CnbError
    from centernet_pytorch_lightning_b200.decode import ctdet_decode
    with pytest.raises(CnbError):
        ctdet_decode(torch.rand(1, 2, 8, 8), torch.rand(1, 2, 8, 8), torch.rand(1, 2, 8, 8))
