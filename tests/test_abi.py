"""The C-ABI library loads and exports every symbol include/lyon_cuda.h declares (no compute without a GPU)."""
import re
from pathlib import Path

import pytest

import oracle_ffi as O
from lyon_b200 import _ffi


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build_cuda()
    L = _ffi.lib()
    header = (O.ROOT / "include" / "lyon_cuda.h").read_text()
    declared = set(re.findall(r"\b(lyon_cuda_\w+)\s*\(", header))
    declared.discard("lyon_cuda_ctx")
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in lyon_cuda.h but not exported"
    assert declared == set(_ffi.SYMBOLS)


def test_struct_layouts_match_the_header():
    import ctypes
    from lyon_b200.geometry import STROKE_VERTEX_DTYPE
    from lyon_b200.options import CStrokeOptions
    assert ctypes.sizeof(CStrokeOptions) == 20
    assert STROKE_VERTEX_DTYPE.itemsize == 32
    assert ctypes.sizeof(_ffi.CPathBatch) == 48


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lyon_b200 import DeviceTessellator
    with pytest.raises(RuntimeError):
        DeviceTessellator()


def test_product_does_not_reference_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/ or tests/emul."""
    for f in (O.ROOT / "lyon_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h", ".cpp"):
            txt = f.read_text()
            for needle in ("liblyon_oracle", "libemul", "oracle_ffi", "emul_ffi", "lyon_oracle.hpp"):
                assert needle not in txt, f"{f} references test infrastructure ({needle})"


def test_wide_vertex_stores_survive_ptxas():
    """The fast-join pass writes each 32-byte vertex with one 256-bit store; ptxas 12.9 was seen to turn the same PTX into
    a 32-bit store in out-of-line code (tools/check_sass.py).  The library that ships must contain the real thing."""
    import shutil
    import subprocess
    import sys
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    root = Path(__file__).resolve().parent.parent
    import __graft_entry__ as g
    lib = g.build_cuda()
    r = subprocess.run([sys.executable, str(root / "tools" / "check_sass.py"), str(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def _c_prototypes(header: str):
    """{name: number of parameters} of every function include/lyon_cuda.h declares."""
    text = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(lyon_cuda_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        params = m.group(2).strip()
        out[m.group(1)] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def test_rust_ffi_declares_the_header():
    """rust/lyon_cuda cannot be compiled in this image (no cargo / rustc), so its extern block is checked mechanically
    against the header: the same functions, each with the same number of parameters."""
    header = (O.ROOT / "include" / "lyon_cuda.h").read_text()
    rust = (O.ROOT / "rust" / "lyon_cuda" / "src" / "ffi.rs").read_text()
    c = _c_prototypes(header)
    assert len(c) >= 30
    r = {}
    for m in re.finditer(r"pub fn (lyon_cuda_\w+)\s*\(([^;]*?)\)\s*(?:->\s*[^;]+)?;", rust, flags=re.S):
        params = m.group(2).strip()
        r[m.group(1)] = 0 if not params else len([p for p in params.split(",") if p.strip()])
    assert set(r) == set(c), f"missing in ffi.rs: {sorted(set(c) - set(r))}; unknown: {sorted(set(r) - set(c))}"
    for name, n in c.items():
        assert r[name] == n, f"{name}: {n} parameters in lyon_cuda.h, {r[name]} in ffi.rs"
    # the structs the two sides share
    assert "pub struct lyon_stroke_vertex" in rust and "pub struct lyon_path_batch" in rust and "pub struct lyon_stroke_options" in rust
