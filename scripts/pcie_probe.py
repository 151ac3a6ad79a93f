"""Pinned host->device bandwidth of a 49 MB query batch, as allocated by default and after pinning the process to the
GPU's NUMA node (nvidia-smi topo / sysfs), to see what the end-to-end figure is bound by."""
import os, subprocess, time
import torch
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
print("cpus allowed:", sorted(os.sched_getaffinity(0))[:64], "of", os.cpu_count())
for node in sorted(os.listdir("/sys/devices/system/node")):
    if node.startswith("node"):
        print(node, open("/sys/devices/system/node/%s/cpulist" % node).read().strip())
def bw(label):
    h = torch.empty((4096, 3000), dtype=torch.float32).pin_memory()
    h.normal_()
    d = torch.empty_like(h, device="cuda")
    o = torch.empty((4096, 100), dtype=torch.float64, device="cuda"); ho = torch.empty((4096, 100), dtype=torch.float64).pin_memory()
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("%s: H2D 49.2 MB in %.3f ms = %.1f GB/s" % (label, ms, 49.152 / ms), flush=True)
bw("default placement")
try:
    bus = torch.cuda.get_device_properties(0).pci_bus_id if hasattr(torch.cuda.get_device_properties(0), "pci_bus_id") else None
    out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip().lower()
    path = "/sys/bus/pci/devices/%s/numa_node" % out[4:] if out.startswith("0000") else "/sys/bus/pci/devices/%s/numa_node" % out
    node = int(open(path).read())
    print("GPU 0", out, "numa node", node)
    if node >= 0:
        cpus = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-"); ids |= set(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            bw("pinned to NUMA node %d cores" % node)
        else:
            print("no allowed cpu on that node")
except Exception as exc:
    print("numa probe failed:", repr(exc))
