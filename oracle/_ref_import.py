"""Import the REAL reference modules (`/root/reference/src/...`) in an image that lacks most of their third-party
dependencies, by stubbing those dependencies.  TEST INFRASTRUCTURE ONLY; used solely by oracle/make_golden.py in the
build container (the GPU box has no /root/reference and never runs this).

The functions we call through this path (tif_image, padded_crop, crop_tif, build_palette,
generate_random_rgb_palette, torch_apply_mask_rgb, SegGptLoss, PromptModel.process_pred_masks, Accumulator.update)
only use numpy / torch / PIL, all of which are real; the stubs exist so that the surrounding `import` lines succeed.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import sys
import types

STUBBED_TOP_LEVEL = (
    "kornia", "lightning", "rasterio", "geopandas", "shapely", "skimage", "matplotlib", "affine", "torchmetrics",
    "omegaconf", "dotenv",
)


class _StubMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_stub(f"{cls.__name__}.{name}")

    def __call__(cls, *a, **k):
        if cls.__dict__.get("_is_base_stub", True) is False:
            return super().__call__(*a, **k)
        return _make_stub(cls.__name__ + "()")


def _make_stub(name: str):
    return _StubMeta(name, (), {"_is_base_stub": True, "__init_subclass__": classmethod(_mark_subclass)})


def _mark_subclass(cls, **kwargs):
    cls._is_base_stub = False


class _StubModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_stub(name)


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUBBED_TOP_LEVEL:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def import_reference(reference_root: str = "/root/reference"):
    """Returns the imported reference modules as a namespace (ml_util, geo_util, model, predict, config)."""
    missing = []
    for top in STUBBED_TOP_LEVEL:
        try:
            importlib.import_module(top)
        except Exception:
            missing.append(top)
    if missing and not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        finder = _StubFinder()
        # only stub what is really missing
        finder.find_spec = (lambda orig: lambda fullname, path, target=None: orig(fullname, path, target)
                            if fullname.split(".")[0] in missing else None)(finder.find_spec)
        sys.meta_path.append(finder)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    ns = types.SimpleNamespace()
    ns.stubbed = tuple(missing)
    ns.config = importlib.import_module("src.config")
    ns.ml_util = importlib.import_module("src.util.ml_util")
    ns.geo_util = importlib.import_module("src.util.geo_util")
    ns.model = importlib.import_module("src.model")
    ns.predict = importlib.import_module("src.predict")
    return ns
