#!/usr/bin/env python3
"""Host -> device copy bandwidth per GPU when 1, 2, 4 ... N ranks copy at the same time (VERDICT r01, weak item 6:
the end-to-end bulk number stops scaling past two GPUs and the limiter is the pinned-host -> device copy chain).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/h2d_scaling.py

Every rank allocates pinned host memory twice -- with the default policy and with the memory policy set to the NUMA
node its GPU hangs off (set_mempolicy(MPOL_PREFERRED), the raw system call: libnuma is not in the image) -- and copies
1 GiB ten times per phase; in phase k only the ranks below k copy.  Rank 0 prints one JSON object."""
import ctypes
import glob
import json
import os
import sys

import torch
import torch.distributed as dist


def read(path, default=""):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return default


def set_mempolicy(mode, node):
    """set_mempolicy(2): mode 1 = MPOL_PREFERRED, 0 = MPOL_DEFAULT.  Returns errno (0 = ok)."""
    libc = ctypes.CDLL(None, use_errno=True)
    SYS_set_mempolicy = 238                     # x86_64
    if mode == 0:
        r = libc.syscall(SYS_set_mempolicy, 0, None, 0)
    else:
        mask = ctypes.c_ulong(1 << node)
        r = libc.syscall(SYS_set_mempolicy, mode, ctypes.byref(mask), 64)
    return 0 if r == 0 else ctypes.get_errno()


def main():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("gloo")
    bus = torch.cuda.get_device_properties(lr).pci_bus_id if hasattr(torch.cuda.get_device_properties(lr), "pci_bus_id") else None
    domain = getattr(torch.cuda.get_device_properties(lr), "pci_domain_id", 0)
    dev = getattr(torch.cuda.get_device_properties(lr), "pci_device_id", 0)
    sysfs = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (domain, bus, dev) if bus is not None else ""
    gpu_node = int(read(sysfs + "/numa_node", "-1") or -1)
    info = {"rank": rank, "pci": sysfs.split("/")[-1], "gpu_numa_node": gpu_node,
            "cpus_allowed": [l.split(":")[1].strip() for l in read("/proc/self/status").splitlines() if l.startswith("Cpus_allowed_list")],
            "mems_allowed": [l.split(":")[1].strip() for l in read("/proc/self/status").splitlines() if l.startswith("Mems_allowed_list")],
            "cpu_now": os.sched_getcpu() if hasattr(os, "sched_getcpu") else -1}
    nbytes = 1 << 30
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    bufs = {}
    bufs["default"] = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    bufs["default"].fill_(1)
    if gpu_node >= 0:
        e = set_mempolicy(1, gpu_node)
        info["set_mempolicy_errno"] = e
        if e == 0:
            b = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            b.fill_(1)                                  # first touch under the policy
            bufs["gpu_local_node"] = b
            set_mempolicy(0, 0)
    # which node did the pages land on?  (numa_maps lists N<node>=pages per mapping)
    def node_of(t):
        addr = "%x" % t.data_ptr()
        for line in read("/proc/self/numa_maps").splitlines():
            if line.startswith(addr[:-3]):
                return [w for w in line.split() if w.startswith("N")]
        return []
    info["pages"] = {k: node_of(v) for k, v in bufs.items()}
    res = {}
    ks = [k for k in (1, 2, 4, 8) if k <= world]
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    for name, src in bufs.items():
        for k in ks:
            for _ in range(2):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            gbs = 0.0
            if rank < k:
                ev0.record()
                for _ in range(10):
                    dst.copy_(src, non_blocking=True)
                ev1.record(); torch.cuda.synchronize()
                gbs = 10 * nbytes / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
            if world > 1:
                dist.barrier()
            res["%s/%d" % (name, k)] = gbs
    info["GBps"] = res
    if world > 1:
        out = [None] * world
        dist.all_gather_object(out, info)
    else:
        out = [info]
    if rank == 0:
        nodes = {os.path.basename(p): read(p + "/cpulist") for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))}
        summary = {}
        for key in out[0]["GBps"]:
            vals = [o["GBps"].get(key, 0.0) for o in out]
            act = [v for v in vals if v > 0]
            summary[key] = {"ranks": len(act), "sum_GBps": round(sum(act), 1), "min_GBps": round(min(act), 1) if act else 0, "max_GBps": round(max(act), 1) if act else 0}
        print(json.dumps({"world": world, "host_numa_nodes": nodes, "nproc": os.cpu_count(), "summary": summary, "ranks": out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
