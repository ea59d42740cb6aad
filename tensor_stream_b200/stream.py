"""TensorStreamConverter -- the reference's Python entry point (tensor_stream/tensor_stream.py:
143-339) in front of the B200 VPP path.

Only the VPP is re-implemented in this repository; demux / NVDEC decode / pacing are out of scope
(SURVEY.md section 2).  So that `read()` can still be exercised end to end, the decoded-frame
source is pluggable and deliberately simple (pull model, one frame per read(), i.e. the
reference's FrameRate.BLOCKING semantics):

  * "synthetic://1920x1080"            seeded random NV12 frames generated on the device
  * a local video file                 software H.264 decode through the FFmpeg libraries bundled
                                       with opencv (tensor_stream_b200/swdecode.py), uploaded per frame
  * any object with next_frame() -> (Y, UV) CUDA uint8 tensors   (e.g. a real NVDEC binding)

The per-frame call below is the reference's getFrame hot section (src/Wrappers/WrapperPython.cpp:
265-350): take a decoded NV12 frame, run VideoProcessor::Convert, hand back a torch tensor.
"""
import logging
from enum import Enum

from .vpp import FourCC, FrameParameters, Planes, ResizeType, VideoProcessor, output_shape  # noqa: F401


class StatusLevel(Enum):  # reference tensor_stream.py:15-21
    OK = 0
    REPEAT = 1
    ERROR = 2


class LogsLevel(Enum):  # :26-34
    NONE = 0
    LOW = 1
    MEDIUM = 2
    HIGH = 3


class LogsType(Enum):  # :39-43
    FILE = 1
    CONSOLE = 2


class FrameRate(Enum):  # :87-97
    NATIVE = 0
    NATIVE_SIMPLE = 1
    NATIVE_LOW_DELAY = 2
    FAST = 3
    BLOCKING = 4


class SyntheticSource:
    """Seeded random NV12 frames created on the device (no decode)."""

    def __init__(self, width, height, device, pitch=None, seed=1234):
        import torch
        self.torch, self.w, self.h, self.dev = torch, width, height, device
        self.pitch = pitch or width
        self.gen = torch.Generator(device=device)
        self.seed, self.index = seed, 0
        self.fps = 25.0

    def next_frame(self):
        t = self.torch
        self.gen.manual_seed(self.seed + self.index)
        y = t.randint(0, 256, (self.h, self.pitch), dtype=t.uint8, device=self.dev, generator=self.gen)
        uv = t.randint(0, 256, (self.h // 2, self.pitch), dtype=t.uint8, device=self.dev, generator=self.gen)
        self.index += 1
        return y, uv


class LocalFileSource:
    """Software decode of a local file (stand-in for the reference's Parser + NVDEC Decoder)."""

    def __init__(self, path, device, max_frames=1 << 30):
        import torch
        from . import swdecode
        self.torch, self.dev = torch, device
        self._frames = iter(swdecode.decode_nv12(path, max_frames) if max_frames < (1 << 20) else
                            swdecode.decode_nv12(path, 4096))
        first = next(self._frames, None)
        if first is None:
            raise RuntimeError("no video frames in %s" % path)
        self._pending = first
        self.h, self.w = first[0].shape
        self.fps = 25.0

    def next_frame(self):
        fr = self._pending if self._pending is not None else next(self._frames, None)
        self._pending = None
        if fr is None:
            raise RuntimeError("Decoding finished")  # what consumers see at end of stream (Decoder.cpp:113-114)
        t = self.torch
        return t.from_numpy(fr[0]).to(self.dev), t.from_numpy(fr[1]).to(self.dev)


class TensorStreamConverter:
    """Same constructor and methods as the reference's class (tensor_stream.py:143-339)."""

    def __init__(self, stream_url, max_consumers=5, cuda_device=0, buffer_size=5, framerate_mode=FrameRate.NATIVE,
                 timeout=None):
        self.log = logging.getLogger(__name__)
        self.stream_url = stream_url
        self.max_consumers = max_consumers
        self.cuda_device = cuda_device
        self.buffer_size = buffer_size
        self.framerate_mode = framerate_mode
        self.timeout = timeout
        self.fps = None
        self.frame_size = None
        self._source = None
        self._vpp = None
        self._index = 0

    # -- configuration calls kept for signature compatibility (logging / NVTX are out of scope) --
    def enable_logs(self, level, log_type):
        self.log.info("logs level %s (%s)", level, log_type)

    def enable_nvtx(self):
        pass

    def skip_analyze(self):
        pass

    def set_timeout(self, value_ms):
        self.timeout = value_ms

    def initialize(self, repeat_number=1):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("Can't initialize TensorStream: no CUDA device (the VPP has no CPU path)")
        dev = torch.device("cuda", self.cuda_device)
        url = self.stream_url
        last = None
        for _ in range(max(1, repeat_number)):  # retry loop of the reference (tensor_stream.py:187-202)
            try:
                if hasattr(url, "next_frame"):
                    self._source = url
                elif isinstance(url, str) and url.startswith("synthetic://"):
                    w, h = (int(v) for v in url[len("synthetic://"):].split("?")[0].lower().split("x"))
                    self._source = SyntheticSource(w, h, dev)
                else:
                    self._source = LocalFileSource(url, dev)
                break
            except (OSError, RuntimeError, ValueError) as e:
                last = e
        if self._source is None:
            raise RuntimeError("Can't initialize TensorStream: %s" % last)
        self.frame_size = (self._source.w, self._source.h)
        self.fps = getattr(self._source, "fps", None)
        self._vpp = VideoProcessor(self.cuda_device)

    def start(self):
        """The reference spawns its producer thread here; this pull-model source needs none."""
        if self._source is None:
            raise RuntimeError("initialize() first")

    def param_read(self, frame_parameters, name="default", delay=0, return_index=False):
        y, uv = self._source.next_frame()
        w, h = self.frame_size
        tensor = self._vpp.convert(y, uv, frame_parameters, width=w, height=h)
        self._index += 1
        return (tensor, self._index) if return_index else tensor

    def read(self, name="default", width=0, height=0, resize_type=ResizeType.NEAREST, crop_coords=(0, 0, 0, 0),
             pixel_format=FourCC.RGB24, planes_pos=Planes.MERGED, normalization=None, delay=0, return_index=False):
        fp = FrameParameters(width=width, height=height, crop_coords=crop_coords, resize_type=resize_type,
                             pixel_format=pixel_format, planes_pos=planes_pos, normalization=normalization)
        return self.param_read(fp, name=name, delay=delay, return_index=return_index)

    def dump(self, tensor, name="default", **_unused):
        """Headerless raw append to <name>.yuv, as the reference's DumpFrame (VideoProcessor.cpp:28-77)."""
        with open(name + ".yuv", "ab") as f:
            f.write(tensor.detach().cpu().contiguous().numpy().tobytes())

    def stop(self):
        if self._vpp is not None:
            self._vpp.close()
            self._vpp = None
        self._source = None
