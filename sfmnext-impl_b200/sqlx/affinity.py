"""CPU / NUMA placement of a rank next to its GPU.

The end-to-end loop is bound by the host -> device copies of the step's inputs from pinned memory (64 MB per step and
rank at the headline workload).  Pinned pages are placed on the NUMA node of the thread that first touches them, and a
rank scheduled on the other socket pays the inter-socket link on every copy: on the round-1 scaling node all eight GPUs
hang off NUMA node 0 while the ranks landed anywhere.  bind_to_gpu() restricts the calling process
to the CPUs local to its GPU's PCIe root complex BEFORE any pinned buffer is allocated, and gives each local rank its
own slice of those CPUs.  Pure host-side plumbing (sysfs + sched_setaffinity); no-op where the topology is unknown.
"""
import os


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(s):
    cpus = []
    for part in (s or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus += list(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_local_cpus(index):
    """(numa_node, [cpu ids]) of the PCIe root complex GPU `index` hangs off, from sysfs; (None, []) when unknown."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(index)
        bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
    except Exception:
        return None, []
    base = "/sys/bus/pci/devices/" + bus
    node = _read(base + "/numa_node")
    cpus = _parse_cpulist(_read(base + "/local_cpulist"))
    try:
        node = int(node) if node is not None else None
    except ValueError:
        node = None
    return node, cpus


def bind_to_gpu(local_rank, local_world=1):
    """Restrict this process to the CPUs local to GPU `local_rank` (a 1/local_world slice of them when several ranks share
    them).  Returns a short description for the bench line."""
    node, cpus = gpu_local_cpus(local_rank)
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return "unsupported platform"
    cpus = [c for c in cpus if c in allowed]
    if not cpus:
        return "no sysfs topology for GPU %d: affinity unchanged (%d CPUs)" % (local_rank, len(allowed))
    # ranks whose GPUs share these CPUs split them evenly (at least 2 CPUs per rank: the main thread + the copy helper)
    share = max(2, len(cpus) // max(1, local_world))
    start = (local_rank * share) % max(1, len(cpus) - share + 1) if len(cpus) > share else 0
    mine = cpus[start:start + share] or cpus
    try:
        os.sched_setaffinity(0, mine)
    except OSError as e:
        return "sched_setaffinity failed: %s" % e
    return "GPU %d: NUMA node %s, process bound to CPUs %s" % (local_rank, node, _fmt(mine))


def _fmt(cpus):
    if not cpus:
        return "-"
    return "%d-%d" % (cpus[0], cpus[-1]) if cpus == list(range(cpus[0], cpus[-1] + 1)) else ",".join(map(str, cpus))
