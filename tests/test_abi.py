"""C-ABI surface: the shared library builds for sm_100a without a GPU, loads, and exports every symbol that
include/mauv_b200.h declares; the ctypes table binds exactly that set. No compute call is made here."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "mauv_b200.h"


def header_symbols():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(mauv_\w+)\s*\(", text))


def test_header_declares_the_hot_path():
    syms = header_symbols()
    for s in ("mauv_sample_weights_f16", "mauv_gemm_f16", "mauv_conv2d_im2col_f16", "mauv_bn_finalize",
              "mauv_sampled_linear_f32", "mauv_mc_reduce", "mauv_kl_fwd_bwd", "mauv_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib_built):
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib_built)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (mauv_\w+)", out))
    missing = header_symbols() - exported
    assert not missing, f"declared in the header but not exported: {sorted(missing)}"


def test_ctypes_table_matches_header(lib_built):
    from mauv import _lib
    assert set(_lib.SIGNATURES) == header_symbols()
    lib = _lib.load()
    assert lib.mauv_version() == 100
    assert lib.mauv_gemm_m_tiles(129) == 2
    assert lib.mauv_kl_chunk_elems() == 4096


def test_sass_is_blackwell_native(lib_built):
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (incl. im2col mode) in the cubin."""
    sass = subprocess.run(["cuobjdump", "-sass", str(lib_built)], capture_output=True, text=True).stdout
    if not sass:
        pytest.skip("cuobjdump unavailable")
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass and "IM2COL" in sass
    assert "HMMA.16816" not in sass  # no legacy mma.sync path


def test_no_cpu_fallback():
    """Product ops refuse CPU tensors instead of silently computing with PyTorch."""
    import torch
    from mauv import _lib
    from mauv.bayesian import Conv2dReparameterization
    layer = Conv2dReparameterization(64, 64, 1, bias=False)
    with pytest.raises(_lib.MauvError):
        layer(torch.zeros(1, 64, 4, 4))


def test_product_does_not_import_oracle():
    pkg = ROOT / "multimodal-auv_b200" / "mauv"
    for f in pkg.rglob("*.py"):
        src = f.read_text()
        assert "bnn_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f
