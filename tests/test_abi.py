"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/ola_gpu.h declares;
the Python mirror binds exactly that set; without a GPU the product path fails loudly (no fallback)."""
import ctypes
import os

import pytest

import olavm_b200
from olavm_b200 import _lib, build


def test_library_is_built_for_sm100a():
    so = build.build()
    assert os.path.exists(so)
    flags = " ".join(build.NVCC_FLAGS)
    assert "arch=compute_100a,code=sm_100a" in flags and "-lineinfo" in flags


def test_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.SO_PATH)
    names = _lib.header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libola_gpu.so does not export {n}"


def test_python_mirror_binds_header_exactly():
    assert sorted(_lib.SIGNATURES) == _lib.header_symbols()
    _lib.load()


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(olavm_b200.OlaError) as e:
        olavm_b200.Context(0)
    assert e.value.code == -1


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.abspath(olavm_b200.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "libola_oracle" not in text, f


def test_rust_ffi_block_is_current():
    """integration/ola_gpu.rs (the extern "C" block a maintainer drops into the reference) is generated from the header."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.call([sys.executable, os.path.join(root, "tools", "gen_rust_ffi.py"), "--check"]) == 0, "run tools/gen_rust_ffi.py"
    rs = open(os.path.join(root, "integration", "ola_gpu.rs")).read()
    from olavm_b200 import _lib

    for name in _lib.header_symbols():
        assert f"pub fn {name}(" in rs, name
