"""jpeg_dsp_studio_b200 - B200-native (sm_100a CUDA) implementation of JPEG-DSP Studio's
compression round trip (reference: engines/pipeline.py::compress_reconstruct).

    from jpeg_dsp_studio_b200 import compress_reconstruct, CompressionParams
    result, intermediate = compress_reconstruct(image_rgb, CompressionParams(quality=50))

Importing this package loads ``libjds.so`` (built by ``python -m
jpeg_dsp_studio_b200.build``); there is no CPU fallback.
"""

from . import _native

_native.load()          # fail loudly at import when the CUDA library is missing

from .models import CompressionParams, CompressionResult, IntermediateData  # noqa: E402
from .engines.pipeline import compress_reconstruct, quality_sweep, compress_batch, compress_stream, plot_payload  # noqa: E402
from .engine import Engine, get_engine  # noqa: E402

__all__ = ['CompressionParams', 'CompressionResult', 'IntermediateData',
           'compress_reconstruct', 'quality_sweep', 'compress_batch', 'compress_stream', 'plot_payload', 'Engine',
           'get_engine']
__version__ = "0.1.0"
