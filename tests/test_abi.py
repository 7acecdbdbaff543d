"""CPU: libbseg.so builds, loads and exports every symbol include/bseg.h declares (no compute without a GPU)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "bseg.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bseg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import beach_seg_b200
    from beach_seg_b200 import _lib

    if not _lib.LIB_PATH.exists():
        beach_seg_b200.build()
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"libbseg.so does not export {n}"
    # the ctypes signature table covers the whole header, nothing more
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_string():
    from beach_seg_b200 import _lib

    L = _lib.lib()
    assert L.bseg_version() >= 1
    assert isinstance(L.bseg_last_error(), bytes)
    assert L.bseg_workspace_bytes(None, 0) == 0
    assert L.bseg_workspace_bytes(None, 1) > 100 * 2**20


def test_no_oracle_import_in_product():
    """The product package must never import the oracle."""
    for p in (ROOT / "beach_seg_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), p


def test_gemm_variant_switch_is_host_state_only():
    """bseg_gemm_set_cta_pairs: a negative argument only queries; the setting round-trips (no GPU involved)."""
    import os

    from beach_seg_b200 import _lib

    L = _lib.lib()
    cur = L.bseg_gemm_set_cta_pairs(-1)
    assert cur in (0, 1)
    if "BSEG_GEMM_2CTA" not in os.environ:
        assert cur == 1  # CTA pairs are the default
    assert L.bseg_gemm_set_cta_pairs(0) == cur
    assert L.bseg_gemm_set_cta_pairs(-1) == 0
    assert L.bseg_gemm_set_cta_pairs(1) == 0
    assert L.bseg_gemm_set_cta_pairs(cur) == 1


def test_train_aug_argument_errors_need_no_gpu():
    """Argument validation of the augmentation entry points happens before any launch."""
    import ctypes as C

    from beach_seg_b200 import _lib

    L = _lib.lib()
    order = (C.c_int32 * 4)(0, 1, 2, 3)
    assert L.bseg_train_aug_fwd(None, None, None, order, None, 0.0, 0.1, _lib.f3((0, 0, 0)), _lib.f3((1, 1, 1)), None,
                                None, None, 0, 8, 8, None) == 0          # empty batch: nothing to do
    assert L.bseg_train_aug_fwd(None, None, None, order, None, 0.0, 0.1, _lib.f3((0, 0, 0)), _lib.f3((1, 1, 1)), None,
                                None, None, 2, 8, 8, None) != 0          # null pointers
    assert b"null argument" in L.bseg_last_error()
    assert L.bseg_train_aug_bwd(None, None, order, _lib.f3((1, 1, 1)), None, None, None, None, 1, 0, 8, None) != 0
    assert b"bad shape" in L.bseg_last_error()
