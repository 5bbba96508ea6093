"""The C-ABI library builds, loads and exports every symbol include/xtag_b200.h declares (no compute calls:
this runs on the CPU build box).  Argument validation that happens before any CUDA call is exercised too."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from xtag_clip_b200 import build, _lib
    build.build(verbose=False)
    return _lib.load()


def _declared():
    text = open(os.path.join(ROOT, "include", "xtag_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(xtag_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from xtag_clip_b200 import _lib
    names = _declared()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/xtag_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table out of sync with the header"
    assert lib.xtag_version() == 1


def test_no_torch_or_cudart_so_dependency():
    """plain C ABI: loadable without torch (static cudart, no libtorch / libc10 in DT_NEEDED)."""
    import subprocess
    from xtag_clip_b200 import _lib
    out = subprocess.run(["readelf", "-d", _lib.LIB_PATH], capture_output=True, text=True).stdout
    needed = re.findall(r"NEEDED.*\[(.*?)\]", out)
    assert not any("torch" in n or "c10" in n for n in needed), needed


def test_argument_validation_without_gpu(lib):
    # null pointers / bad sizes are rejected before any CUDA work and leave a message
    rc = lib.xtag_clip_fwd(None, None, 1, 4, 4, 8, None, 0, None, None, None, None, 0, 0, None)
    assert rc == -1 and b"null" in lib.xtag_last_error()
    rc = lib.xtag_lse_combine(None, 2, 8, None, None)
    assert rc == -1
    assert lib.xtag_clip_fwd_ws_bytes(0, 4, 8, 1, 0) == 0
    assert lib.xtag_clip_fwd_ws_bytes(4096, 4096, 512, 1, 0) > 0
    assert lib.xtag_clip_bwd_ws_bytes(4096, 4096, 512, 1, 2) >= 4096 * 4096 * 2      # bf16 dS staging
    # tcgen05 path refuses shapes TMA cannot describe
    assert lib.xtag_clip_fwd_ws_bytes(64, 64, 12, 1, 2) == 0


def test_product_path_fails_loudly_without_library(monkeypatch, tmp_path):
    from xtag_clip_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
