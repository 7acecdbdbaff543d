"""beach_seg_b200 — B200 (sm_100a) implementation of beach_seg's segmentation hot path.

Host side mirrors the reference's Python interface (src/model.py, src/util/ml_util.py, src/predict*.py); the
arithmetic runs in hand-written CUDA kernels behind the C ABI of libbseg.so (include/bseg.h)."""
from ._lib import BsegError, build, lib  # noqa: F401

__all__ = ["BsegError", "build", "lib"]
