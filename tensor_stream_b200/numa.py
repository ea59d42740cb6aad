"""Host-memory placement for the zero-copy host path.

`tsvpp_convert_batch_host` lets the kernel read pinned host frames in place over PCIe.  Those reads
are SM-issued and latency-bound: from most pinned allocations they run at PCIe speed (29-30 k 1080p
frames/s on a B200), from some at a quarter of that (7 k) -- whole-frame cudaMemcpy (copy engines)
hardly notices.  The slow case was first blamed on buffers living on the other socket, but it has
since been seen with buffers bound to the GPU's own node too, so treat placement as something to
probe, not to assume (bench.py does: GPU node first, then the default placement, then the other
node, keeping the fastest).  Pinned buffers land on the node of the thread that allocates them:

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


def all_nodes(sysfs="/sys"):
    """NUMA nodes that have CPUs, ascending."""
    out = []
    try:
        names = os.listdir(os.path.join(sysfs, "devices/system/node"))
    except OSError:
        return out
    for name in names:
        if name.startswith("node") and name[4:].isdigit() and node_cpus(int(name[4:]), sysfs):
            out.append(int(name[4:]))
    return sorted(out)


@contextlib.contextmanager
def bound_to_node(node, sysfs="/sys"):
    """Run the body with the calling thread restricted to the CPUs of NUMA node `node` (no-op when the
    node has no CPU this process may use).  Yields the node or None."""
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
