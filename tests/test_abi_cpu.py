"""The C-ABI library builds, loads without a GPU and exports every symbol include/mfm_b200.h declares.
No compute call is made here."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    from factorized_b200 import cuda_ops
    if not os.path.exists(cuda_ops.LIB_PATH):
        ge.build()
    return cuda_ops.load_library()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mfm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mfm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libmfm_b200.so does not export %s" % n


def test_binding_covers_header(lib):
    from factorized_b200 import cuda_ops
    assert sorted(cuda_ops.EXPORTS) == declared_symbols()
    assert lib.mfm_version() == 100
    assert lib.mfm_get_gemm_path() in (0, 1, 2)


def test_product_fails_loudly_without_library_or_gpu(tmp_path):
    import torch
    from factorized_b200 import cuda_ops, MFM
    with pytest.raises(RuntimeError):
        cuda_ops.load_library(str(tmp_path / "nope.so"))
    from oracle import mfm_oracle as O
    m = MFM(*O.tiny_configs())
    with pytest.raises(RuntimeError):
        m.forward(torch.zeros(4, 6, 15))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            cuda_ops.CudaOps()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "factorized_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            s = open(os.path.join(pkg, f)).read()
            assert "oracle" not in s.replace("mfm_oracle", "oracle") or "import oracle" not in s
            assert not re.search(r"^\s*(from|import)\s+(oracle|tests|emu_ops)", s, flags=re.M), f
