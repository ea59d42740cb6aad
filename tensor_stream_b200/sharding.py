"""Multi-GPU layout of the VPP path: none of it touches pixels.

Frames / streams are independent (reference python_examples/different_streams.py:79-99 runs one
TensorStream instance per stream per device), so the path shards with NO data-path
collective: stream s -> rank s mod world, each rank owns a VideoProcessor + frame pool on its own
GPU.  The only cross-rank traffic is the benchmark's timing reduction (max over ranks).
"""


def streams_of_rank(num_streams, world_size, rank):
    """Streams (or frame batches) owned by `rank`: s mod world_size == rank."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world")
    return [s for s in range(num_streams) if s % world_size == rank]


def rank_of_stream(stream_id, world_size):
    return stream_id % world_size


def max_over_ranks(values, device=None, group=None):
    """Element-wise MAX of a list of floats over all ranks (device-timed milliseconds in
    bench.py).  Works on any backend (nccl on GPUs, gloo in the CPU tests); identity when
    torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return [float(x) for x in t.cpu()]


def aggregate_fps(frames_per_rank_per_step, steps, max_ms, world_size):
    """Whole-job throughput: all ranks' frames over the slowest rank's time (weak scaling)."""
    return world_size * frames_per_rank_per_step * steps / (max_ms * 1e-3)
