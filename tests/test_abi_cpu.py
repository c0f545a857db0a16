"""CPU-side checks of the C-ABI library: it builds, loads, exports every symbol the header
declares, and its host-side entry points / argument validation behave (no kernels launched)."""
import ctypes
import os
import re

import numpy as np
import pytest

from moma_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _lib.load()


def header_functions():
    text = open(os.path.join(ROOT, "include", "moma_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(moma_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    declared = header_functions()
    assert declared, "no functions parsed from the header"
    assert sorted(_lib.SIGNATURES) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/moma_b200.h but not exported"


def test_no_stray_exports():
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", _build.LIB_PATH], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    ours = {e for e in exported if e.startswith("moma_")}
    assert ours == set(header_functions())


def test_version_and_error_string(lib):
    assert lib.moma_abi_version() == _lib.ABI_VERSION
    # validation failure -> negative code + message, no launch attempted
    rc = lib.moma_l2norm_fwd(None, None, 4, 6, 1e-12, None)        # D % 4 != 0
    assert rc == -3 and b"multiple of 4" in lib.moma_last_error()
    rc = lib.moma_enqueue(16, 8, 8, 16, None, 4, 0, None, 0, 1, 0, 1e-12, None)   # n > K
    assert rc == -1 and b"duplicate ids" in lib.moma_last_error()
    rc = lib.moma_enqueue(16, 2, 8, 16, None, 9, 0, None, 0, 2, 0, 1e-12, None)   # K % world != 0
    assert rc == -1
    rc = lib.moma_attn_fwd(16, 16, None, 16, 16, 8, 40, 4, 16, 16, 16, 16, None, None, None)   # head_dim 10
    assert rc == -3 and b"head_dim" in lib.moma_last_error()
    rc = lib.moma_nce_partial(16, 16, 8, 1024, 64, 1.0, _lib.F32, 1, 16, 16, 16, 16, None)   # D > 512
    assert rc == -3


def test_ema_plan_host_side(lib):
    """The EMA chunk table is host logic: sizes, pointers, per-chunk counts."""
    numels = [7, 9408, 64, 8192, 8193, 0, 20000]
    n = len(numels)
    arr = (ctypes.c_int64 * n)(*numels)
    chunks, nbytes = ctypes.c_int64(0), ctypes.c_size_t(0)
    assert lib.moma_ema_plan_size(n, arr, ctypes.byref(chunks), ctypes.byref(nbytes)) == 0
    want = sum((x + 8191) // 8192 for x in numels)
    assert chunks.value == want and nbytes.value == want * 32
    base_s, base_d = 0x10000000, 0x20000000
    sp = (ctypes.c_void_p * n)(*[base_s + 0x100000 * i for i in range(n)])
    dp = (ctypes.c_void_p * n)(*[base_d + 0x100000 * i + (4 if i == 2 else 0) for i in range(n)])
    table = np.zeros(nbytes.value, dtype=np.uint8)
    assert lib.moma_ema_plan_fill(n, sp, dp, arr, table.ctypes.data, table.size) == 0
    rec = table.view(np.dtype([("src", "<u8"), ("dst", "<u8"), ("count", "<i4"), ("vec", "<i4"), ("pad", "<i8")]))
    assert rec["count"].sum() == sum(numels)
    assert rec["count"].max() <= 8192
    assert rec["src"][0] == base_s and rec["dst"][0] == base_d and rec["count"][0] == 7
    # tensor 2 has a destination that is only 4-byte aligned -> scalar path
    i2 = 1 + 2   # chunk index: tensor0 (1 chunk) + tensor1 (2 chunks)
    assert rec["vec"][i2] == 0 and rec["vec"][0] == 1
    # tensor 6 (20000 elements): 3 chunks, contiguous
    last = rec[-3:]
    assert list(last["count"]) == [8192, 8192, 20000 - 16384]
    assert last["src"][1] - last["src"][0] == 8192 * 4
    # too-small table is refused
    assert lib.moma_ema_plan_fill(n, sp, dp, arr, table.ctypes.data, 32) == -5


def test_splits_heuristic_is_host_side(lib):
    s = lib.moma_nce_num_splits(512, 128, 65536, _lib.F32)
    assert 1 <= s <= 1024
    assert lib.moma_nce_num_splits(32, 128, 64, _lib.F32) == 1
    assert lib.moma_attn_bwd_workspace_bytes(512, 128, 4) >= (512 * 128 * 4 + 4 * 512) * 4


def test_linear_and_peer_validation_is_host_side(lib):
    """Argument checks of the newer entry points return before any CUDA call (no GPU here)."""
    assert lib.moma_linear_fwd(None, None, None, 4, 4, 4, 0, None, None, 0, None) == -1
    assert b"null pointer" in lib.moma_last_error()
    assert lib.moma_linear_fwd(16, 16, None, 0, 4, 4, 0, 16, None, 0, None) == -1 and b"bad shape" in lib.moma_last_error()
    assert lib.moma_linear_bwd(16, 16, None, 16, 4, 4, 4, 1, None, None, None, None, 0, None) == -1   # relu without y
    # split-K workspace: a 256 x 128 output over K = 512 is split (32 tiles on 148 SMs): room for >= 2 partial tiles
    assert lib.moma_linear_workspace_bytes(256, 128, 512) > 2 * 256 * 128 * 4
    assert lib.moma_linear_workspace_bytes(0, 1, 1) == 0
    ctrl = lib.moma_peer_ctrl_bytes()
    assert ctrl >= 48 and ctrl % 256 == 0
    args = dict(src=16, stride=0, nbytes=256, cast=0, bases=16, ctrl_off=0, data_off=256, region=1 << 20, rank=0, world=2, ch=0, out=16)

    def call(**kw):
        a = dict(args, **kw)
        return lib.moma_peer_exchange(a["src"], a["stride"], a["nbytes"], a["cast"], a["bases"], a["ctrl_off"], a["data_off"],
                                      a["region"], a["rank"], a["world"], a["ch"], a["out"], None)
    assert call(world=17) == -1 and b"max world" in lib.moma_last_error()
    assert call(rank=2) == -1
    assert call(ch=4) == -1 and b"channel" in lib.moma_last_error()
    assert call(nbytes=100) == -2 and b"multiples of 16" in lib.moma_last_error()
    assert call(region=512) == -5 and b"region too small" in lib.moma_last_error()       # needs 2 x world x bytes
    assert call(src=None) == -1


def test_peer_exchange_unavailable_without_cuda():
    import torch
    from moma_b200.peer import PeerExchange
    assert PeerExchange.create(None, torch.device("cpu"), 1 << 20) is None


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_build, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
