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


def test_shim_exports_the_reference_modules_public_names():
    """The repo-root ``mfm_model`` must resolve the reference scripts' import lines unchanged (mfm_mosi.py:30-31, mfm_moud.py,
    mfm_you.py, mfm_mmmo.py import from ``mfm_model``): every class and module-level function of /root/reference/mfm_model.py."""
    import importlib
    shim = importlib.import_module("mfm_model")
    names = ["compute_kernel", "loss_MMD", "loss_KLD", "encoderLSTM", "decoderLSTM", "MFN", "M_A", "M_B", "M_C", "M_D", "MFM",
             "MFM_KL_EF", "MFM_KL", "MFM_missing", "seq2seq", "basic_missing"]
    missing = [n for n in names if not hasattr(shim, n)]
    assert not missing, missing
    import factorized_b200 as F
    for fn in ("train_mfm", "train_mfm_ablation", "train_mfm_missing", "train_mfm_test_zeros"):
        assert callable(getattr(F, fn)), fn


def test_score_block():
    """train.score restates mfm_mosi.py:483-498 as a dict."""
    import numpy as np
    from factorized_b200.train import score
    y = np.array([-2.2, -0.4, 0.3, 1.6, 2.7, -1.1])
    p = np.array([-1.8, 0.2, 0.4, 1.4, 1.2, -0.9])
    s = score(p, y)
    assert abs(s["mae"] - float(np.mean(np.abs(p - y)))) < 1e-12
    assert abs(s["corr"] - float(np.corrcoef(p, y)[0][1])) < 1e-12
    assert s["mult_acc"] == round(float(np.mean(np.round(p) == np.round(y))), 5)
    assert abs(s["binary_acc"] - 5.0 / 6.0) < 1e-12
    if "confusion_matrix" in s:
        assert s["confusion_matrix"] == [[2, 1], [0, 3]] and 0.0 <= s["mult_f_score"] <= 1.0
    c = score(np.array([[0.1, 0.9], [0.8, 0.2], [0.3, 0.7]]), np.array([1, 1, 0]), head="ce")      # mfm_moud.py:421-428
    assert abs(c["acc"] - 1.0 / 3.0) < 1e-12
    if "confusion_matrix" in c:
        assert c["confusion_matrix"] == [[0, 1], [1, 1]] and "accuracy" in c["classification_report"]


def test_header_is_plain_c_and_links_from_a_c_program(lib, tmp_path):
    """The drop-in boundary is a C ABI: include/mfm_b200.h must compile as C99 (no C++, no torch types) and a C program must
    link against the library and call it.  ``mfm_version`` and an argument-check error path only: no GPU work."""
    import shutil
    import subprocess
    from factorized_b200 import cuda_ops
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include "mfm_b200.h"\n'
                   'int main(void) {\n'
                   '  int rc = mfm_zero(0, 0, (void*)0);           /* bad argument: must return an error code, not crash */\n'
                   '  printf("%d %d\\n", mfm_version(), rc);\n  return 0;\n}\n')
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, "-fsyntax-only", str(src)], check=True)
    libdir = os.path.dirname(cuda_ops.LIB_PATH)
    exe = tmp_path / "t"
    subprocess.run(["gcc", "-std=c99", "-I", inc, str(src), "-o", str(exe), "-L", libdir, "-lmfm_b200",
                    "-Wl,-rpath," + libdir], check=True)
    env = dict(os.environ)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, env=env).stdout.split()
    assert out[0] == "100" and int(out[1]) != 0
