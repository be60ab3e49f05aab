"""Host placement for one-process-per-GPU jobs: run the process (and allocate its pinned staging buffers, which follow
the first-touch policy) on the NUMA node the GPU hangs off.  On a two-socket 8-GPU box, host buffers on the far socket
make every H2D / D2H copy cross the socket interconnect; with all ranks copying at once that link, not PCIe, bounds
the end-to-end rate.  No effect (and no error) when sysfs does not describe the topology."""
import os


def gpu_numa_node(pci_bus_id):
    """NUMA node of a PCI device ("0000:1b:00.0", case-insensitive), or None."""
    for cand in (pci_bus_id.lower(), pci_bus_id.lower()[-12:], "0000:" + pci_bus_id.lower()[-7:]):
        try:
            with open(f"/sys/bus/pci/devices/{cand}/numa_node") as f:
                node = int(f.read().strip())
            return node if node >= 0 else None
        except (OSError, ValueError):
            continue
    return None


def node_cpus(node):
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            text = f.read().strip()
    except OSError:
        return set()
    cpus = set()
    for part in text.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_node(device_index):
    """Restrict this process to the CPUs of the NUMA node of CUDA device `device_index`.  Returns a dict describing
    what was done ({"node": n, "cpus": count} or {"node": None, "why": ...})."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
    except Exception as e:          # old torch without the pci_* properties
        return {"node": None, "why": f"no PCI id ({type(e).__name__})"}
    node = gpu_numa_node(bus)
    if node is None:
        return {"node": None, "why": f"sysfs has no numa_node for {bus}"}
    cpus = node_cpus(node) & os.sched_getaffinity(0)
    if not cpus:
        return {"node": node, "why": "no allowed CPU on that node"}
    os.sched_setaffinity(0, cpus)
    return {"node": node, "cpus": len(cpus), "pci": bus}
