"""Import the UNMODIFIED reference from /root/reference (TEST INFRASTRUCTURE).

Works where ``/root/reference`` exists (the build container) or where ``oracle/make_ref.py``
has staged its hot-path files under ``oracle/_ref/reference`` (the GPU box: bench.py's
``--impl reference`` / ``cpu_baseline`` legs only - the ``-m gpu`` tests never read either).  The reference's ``utils/metrics.py:5`` imports ``skimage.metrics`` at module
top, so ``oracle.skimage_standin`` is registered first.  The reference's top-level
package names (``engines``, ``models``, ``utils``) are removed from ``sys.modules``
again after loading so they cannot shadow this repo's drop-in packages of the same
names; the returned namespace keeps the loaded modules alive.
"""

import importlib
import importlib.util
import os
import sys
import types

def _default_root():
    """/root/reference in the build container, else the copy oracle/make_ref.py staged under
    oracle/_ref/reference (git-ignored; it travels to the GPU box with the snapshot)."""
    if os.path.isfile("/root/reference/engines/pipeline.py"):
        return "/root/reference"
    staged = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")
    return staged if os.path.isfile(os.path.join(staged, "engines", "pipeline.py")) else "/root/reference"


REFERENCE_ROOT = os.environ.get("JDS_REFERENCE_ROOT") or _default_root()
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


_alias_cache = None


def load_aliasing_demo() -> types.SimpleNamespace:
    """The reference's aliasing-demo callables (gui/dialogs/aliasing_demo_dialog.py) -
    pattern generators, ``compute_metrics`` and ``AliasingDemoWorker`` - imported from the
    UNMODIFIED file.  PySide6 is not installed, so inert stand-ins for the few Qt names the
    module touches at import time are registered first: ``QObject`` becomes ``object`` and
    ``Signal`` an object whose ``emit`` stores its payload, which is all the worker's
    ``run()`` needs.  ``run_worker(image, quality)`` returns the dict the worker emits."""
    global _alias_cache
    if _alias_cache is not None:
        return _alias_cache
    ref = load()
    import cv2  # noqa: F401  (the module needs it)

    class _Signal:
        def __init__(self, *a, **k):
            self.payload = []

        def emit(self, *args):
            self.payload.append(args)

        def connect(self, *a, **k):
            pass

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

    def _qt_module(name):
        m = types.ModuleType(name)
        m.__getattr__ = lambda attr: type(attr, (_Anything,), {})
        return m

    stubs = {n: _qt_module(n) for n in ("PySide6", "PySide6.QtWidgets", "PySide6.QtCore", "PySide6.QtGui")}
    stubs["PySide6.QtCore"].QObject = object
    stubs["PySide6.QtCore"].Signal = _Signal
    stubs["PySide6.QtCore"].QThread = _Anything
    stubs["PySide6.QtCore"].Qt = _Anything()
    combobox = types.ModuleType("gui.widgets.styled_combobox")
    combobox.style_combobox = lambda *a, **k: None
    pkg_gui, pkg_widgets = types.ModuleType("gui"), types.ModuleType("gui.widgets")
    pkg_gui.__path__, pkg_widgets.__path__ = [], []
    injected = dict(stubs)
    injected.update({"gui": pkg_gui, "gui.widgets": pkg_widgets, "gui.widgets.styled_combobox": combobox})
    # the reference's own packages, as the module expects to import them
    for k, v in ref._modules.items():
        injected[k] = v
    saved = {k: sys.modules.get(k) for k in injected}
    sys.modules.update(injected)
    try:
        path = os.path.join(REFERENCE_ROOT, "gui", "dialogs", "aliasing_demo_dialog.py")
        spec = importlib.util.spec_from_file_location("_ref_aliasing_demo_dialog", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    def run_worker(image, quality=50):
        w = mod.AliasingDemoWorker(image, quality)
        # class-level Signal objects are shared between instances: give this one its own
        w.finished, w.progress, w.error = _Signal(), _Signal(), _Signal()
        w.run()
        if w.error.payload:
            raise RuntimeError(w.error.payload[0][0])
        return w.finished.payload[0][0]

    _alias_cache = types.SimpleNamespace(
        module=mod, run_worker=run_worker, compute_metrics=mod.compute_metrics,
        generate_equiluminance_stripes=mod.generate_equiluminance_stripes,
        generate_chroma_checkerboard=mod.generate_chroma_checkerboard,
        generate_1px_checkerboard=mod.generate_1px_checkerboard)
    return _alias_cache
