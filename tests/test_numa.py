"""CPU: host-placement helper (rag4dyg_b200/numa.py) parses sysfs cpulists and degrades to a no-op when the topology is
not described (containers report numa_node = -1 or nothing at all)."""
import os

from rag4dyg_b200 import numa


def test_node_cpus_and_missing_nodes():
    cpus = numa.node_cpus(0)
    assert isinstance(cpus, set) and (not cpus or all(isinstance(c, int) for c in cpus))
    assert numa.node_cpus(10_000) == set()
    assert numa.gpu_numa_node("ffff:ff:1f.0") is None            # no such PCI device


def test_bind_is_a_noop_without_topology(monkeypatch):
    before = os.sched_getaffinity(0)
    info = numa.bind_to_gpu_node(0)                                # no GPU / no numa_node here: must not raise
    assert info["node"] is None or "cpus" in info or "why" in info
    if info["node"] is None:
        assert os.sched_getaffinity(0) == before
