"""Host-memory placement for the zero-copy host path.

`tsvpp_convert_batch_host` lets the kernel read pinned host frames in place over PCIe.  Those reads
are latency-bound SM-issued requests: from the GPU's own NUMA node they run at PCIe speed (29.7 k
1080p frames/s on a B200), from the other socket at a quarter of that (6.8 k, measured on the same
box type) -- whole-frame cudaMemcpy (copy engines) hardly notices the difference.  Pinned buffers land
on the node of the thread that allocates them, so allocate them with the thread bound to the GPU's
node:

    with numa.bound_to_gpu(0):
        frames = [torch.empty(...).pin_memory() for _ in range(n)]
"""
import contextlib
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_pci_bus_id(index):
    """'0000:d1:00.0' of CUDA device `index`, or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        return "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except Exception:
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        b = pynvml.nvmlDeviceGetPciInfo(h).busId
        b = b.decode() if isinstance(b, bytes) else b
        return b[-12:].lower()
    except Exception:
        return None


def gpu_numa_node(index, sysfs="/sys"):
    bus = gpu_pci_bus_id(index)
    if not bus:
        return None
    try:
        node = int(open(os.path.join(sysfs, "bus/pci/devices", bus, "numa_node")).read())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def node_cpus(node, sysfs="/sys"):
    try:
        return _parse_cpulist(open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)).read())
    except OSError:
        return set()


@contextlib.contextmanager
def bound_to_gpu(index, sysfs="/sys"):
    """Run the body with the calling thread restricted to the CPUs of the GPU's NUMA node (no-op when
    the topology cannot be read or the node has no CPU this process may use).  Yields the node or None."""
    node = gpu_numa_node(index, sysfs)
    old = os.sched_getaffinity(0)
    cpus = (node_cpus(node, sysfs) & old) if node is not None else set()
    if not cpus:
        yield None
        return
    os.sched_setaffinity(0, cpus)
    try:
        yield node
    finally:
        os.sched_setaffinity(0, old)
