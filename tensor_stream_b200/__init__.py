"""tensor_stream_b200 -- B200-native (sm_100a) video-post-processing path of osai-ai/tensor-stream.

    from tensor_stream_b200 import VideoProcessor, FrameParameters, FourCC, Planes, ResizeType
    from tensor_stream_b200 import TensorStreamConverter        # the reference's Python entry point

The native library (tensor_stream_b200/libtsvpp.so) is loaded on first use; there is no CPU fallback.
"""
from .vpp import (FourCC, FrameParameters, Planes, ResizeType, VideoProcessor, VppError,  # noqa: F401
                  convert_batch_aten, load_library, output_info, output_shape)
from . import numa  # noqa: F401  (host-memory placement for the zero-copy host path)
from .stream import FrameRate, LogsLevel, LogsType, StatusLevel, TensorStreamConverter  # noqa: F401
