"""Import the UNMODIFIED reference from /root/reference (TEST INFRASTRUCTURE).

Only works where ``/root/reference`` exists (the build container; never the GPU
box).  The reference's ``utils/metrics.py:5`` imports ``skimage.metrics`` at module
top, so ``oracle.skimage_standin`` is registered first.  The reference's top-level
package names (``engines``, ``models``, ``utils``) are removed from ``sys.modules``
again after loading so they cannot shadow this repo's drop-in packages of the same
names; the returned namespace keeps the loaded modules alive.
"""

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("JDS_REFERENCE_ROOT", "/root/reference")
_TOP = ("engines", "models", "utils")
_cache = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "engines", "pipeline.py"))


def load() -> types.SimpleNamespace:
    """Return a namespace with the reference's own callables."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    from . import skimage_standin
    skimage_standin.install()

    saved = {k: v for k, v in sys.modules.items()
             if k.split(".")[0] in _TOP}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        pipeline = importlib.import_module("engines.pipeline")
        color_space = importlib.import_module("engines.color_space")
        dct_engine = importlib.import_module("engines.dct_engine")
        quantizer = importlib.import_module("engines.quantizer")
        block_processor = importlib.import_module("engines.block_processor")
        metrics = importlib.import_module("utils.metrics")
        constants = importlib.import_module("utils.constants")
        test_images = importlib.import_module("utils.test_images")
        models = importlib.import_module("models")
        for m in (pipeline, color_space, dct_engine, quantizer, block_processor,
                  metrics, constants, test_images, models):
            assert os.path.realpath(m.__file__).startswith(
                os.path.realpath(REFERENCE_ROOT)), m.__file__
    finally:
        sys.path.remove(REFERENCE_ROOT)
        loaded = {k: v for k, v in sys.modules.items()
                  if k.split(".")[0] in _TOP}
        for k in loaded:
            del sys.modules[k]
        sys.modules.update(saved)

    _cache = types.SimpleNamespace(
        compress_reconstruct=pipeline.compress_reconstruct,
        CompressionParams=models.CompressionParams,
        CompressionResult=models.CompressionResult,
        IntermediateData=models.IntermediateData,
        pipeline=pipeline, color_space=color_space, dct_engine=dct_engine,
        quantizer=quantizer, block_processor=block_processor, metrics=metrics,
        constants=constants, test_images=test_images, _modules=loaded,
    )
    return _cache
