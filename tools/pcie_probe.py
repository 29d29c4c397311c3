"""Pinned-memory H2D / D2H bandwidth of cuda:0 with the default CPU affinity and with the process bound to the
GPU's NUMA node before the pinned buffers are allocated (first touch decides where the pages live)."""
import os
import time

import torch


def gpu_numa(dev=0):
    p = torch.cuda.get_device_properties(dev)
    bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    base = f"/sys/bus/pci/devices/{bdf}"
    node = open(base + "/numa_node").read().strip() if os.path.exists(base + "/numa_node") else "?"
    cpus = open(base + "/local_cpulist").read().strip() if os.path.exists(base + "/local_cpulist") else ""
    return bdf, node, cpus


def parse_cpulist(s):
    out = set()
    for part in s.split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out


def bw(nbytes=1 << 30, reps=5):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda:0")
    res = {}
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        res[name] = round(nbytes * reps / (time.perf_counter() - t) / 1e9, 1)
    return res


if __name__ == "__main__":
    torch.cuda.init()
    bdf, node, cpus = gpu_numa()
    print("gpu", bdf, "numa_node", node, "local_cpulist", cpus, "affinity", len(os.sched_getaffinity(0)), "cpus")
    print("nodes online:", open("/sys/devices/system/node/online").read().strip() if os.path.exists("/sys/devices/system/node/online") else "?")
    print("default affinity :", bw())
    local = parse_cpulist(cpus) & os.sched_getaffinity(0)
    if local:
        os.sched_setaffinity(0, local)
        print("bound to GPU node:", bw(), len(local), "cpus")
    else:
        print("no local cpu list available inside this container")
