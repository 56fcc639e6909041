"""Bind a rank's host threads (and therefore its first-touch pinned buffers) to the NUMA node its GPU hangs off.

An 8-GPU B200 node has two CPU sockets; a process that pins its staging buffers on the other socket pays the inter-socket
link on every host<->device copy, and eight ranks doing so share that link.  `bind_to_gpu_node(i)` reads the GPU's PCI
bus id from the CUDA runtime, its NUMA node from sysfs and restricts the process to that node's CPUs BEFORE the pinned
buffers are allocated (Linux places pages on the node of the thread that first touches them).  Best effort: returns a
description of what it did, never raises."""
import os


def _cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            out.extend(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def gpu_numa_node(device_index):
    import torch
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_node(device_index):
    info = {"numa_node": None, "bound": False, "cpus": None}
    node = gpu_numa_node(device_index)
    info["numa_node"] = node
    if node is None:
        return info
    try:
        cpus = _cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"], info["cpus"] = True, len(allowed)
    except Exception as e:
        info["error"] = str(e)[:80]
    return info
