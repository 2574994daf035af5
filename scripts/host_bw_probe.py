#!/usr/bin/env python
"""Where does the host side of a multi-GPU box saturate?  Concurrent pinned H2D + D2H copies on every GPU, with the
pinned buffers (a) allocated the default way by one thread and (b) bound to the NUMA node of the GPU they feed
(mmap + mbind + cudaHostRegister).  Prints the topology the container can see and one JSON line per experiment.

Usage: python scripts/host_bw_probe.py [MiB per buffer, default 512]"""
import ctypes
import glob
import json
import mmap
import os
import subprocess
import sys
import time

import torch

MIB = int(sys.argv[1]) if len(sys.argv) > 1 else 512
NBYTES = MIB << 20


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return f"<{e}>"


def gpu_numa_nodes():
    nodes = []
    for g in range(torch.cuda.device_count()):
        bus = sh(f"nvidia-smi -i {g} --query-gpu=pci.bus_id --format=csv,noheader").strip()
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        path = f"/sys/bus/pci/devices/{bus}/numa_node"
        try:
            nodes.append(int(open(path).read()))
        except Exception:  # noqa: BLE001
            nodes.append(-1)
    return nodes


libc = ctypes.CDLL(None, use_errno=True)
SYS_MBIND = 237  # x86-64
MPOL_BIND, MPOL_PREFERRED = 2, 1


def numa_buffer(nbytes, node):
    """Anonymous mapping bound to `node`, touched, then page-locked for CUDA.  Returns (uint8 tensor, keepalive)."""
    mm = mmap.mmap(-1, nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    buf = (ctypes.c_uint8 * nbytes).from_buffer(mm)
    addr = ctypes.addressof(buf)
    status = "default"
    if node >= 0:
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(SYS_MBIND, ctypes.c_void_p(addr), ctypes.c_ulong(nbytes), MPOL_BIND, ctypes.byref(mask),
                          ctypes.c_ulong(64), 0)
        status = "bound" if rc == 0 else f"mbind errno {ctypes.get_errno()}"
    ctypes.memset(addr, 1, nbytes)  # first touch under the policy
    t = torch.frombuffer(buf, dtype=torch.uint8)
    rc = torch.cuda.cudart().cudaHostRegister(addr, nbytes, 0)
    if int(rc) != 0:
        status += f" cudaHostRegister rc {int(rc)}"
    return t, (mm, buf, addr), status


def run(name, srcs, dsts, devs, which):
    """Concurrent copies on the GPUs in `devs`: which in {"h2d", "d2h", "both"}; returns aggregate GB/s."""
    streams = {g: (torch.cuda.Stream(g), torch.cuda.Stream(g)) for g in devs}
    dbuf = {g: (torch.empty(NBYTES, dtype=torch.uint8, device=f"cuda:{g}"),
                torch.empty(NBYTES, dtype=torch.uint8, device=f"cuda:{g}")) for g in devs}

    def issue():
        for g in devs:
            if which in ("h2d", "both"):
                with torch.cuda.stream(streams[g][0]):
                    dbuf[g][0].copy_(srcs[g], non_blocking=True)
            if which in ("d2h", "both"):
                with torch.cuda.stream(streams[g][1]):
                    dsts[g].copy_(dbuf[g][1], non_blocking=True)

    def sync():
        for g in devs:
            torch.cuda.synchronize(g)
    issue()
    sync()
    reps = 4
    t0 = time.perf_counter()
    for _ in range(reps):
        issue()
    sync()
    dt = time.perf_counter() - t0
    per = (2 if which == "both" else 1) * NBYTES * len(devs) * reps
    print(json.dumps({"experiment": name, "gpus": list(devs), "direction": which, "GB_per_s": per / dt / 1e9}), flush=True)


def main():
    n = torch.cuda.device_count()
    print("== topology")
    print(sh("nvidia-smi topo -m"))
    print("nproc", sh("nproc"), "| allowed cpus", sh("grep Cpus_allowed_list /proc/self/status"),
          "| allowed mems", sh("grep Mems_allowed_list /proc/self/status"))
    for p in sorted(glob.glob("/sys/devices/system/node/node*/cpulist")):
        print(p, open(p).read().strip(), "| MemTotal", sh(f"grep MemTotal {os.path.dirname(p)}/meminfo"))
    nodes = gpu_numa_nodes()
    print("gpu -> numa node", nodes)
    devs = list(range(n))
    # (a) default pinned allocation (one thread, the CUDA allocator)
    srcs = {g: torch.empty(NBYTES, dtype=torch.uint8).pin_memory() for g in devs}
    dsts = {g: torch.empty(NBYTES, dtype=torch.uint8).pin_memory() for g in devs}
    for g in (0, n - 1):
        for which in ("h2d", "d2h", "both"):
            run("default-alloc single gpu", srcs, dsts, [g], which)
    for which in ("h2d", "d2h", "both"):
        run("default-alloc all gpus", srcs, dsts, devs, which)
    if n >= 4:
        run("default-alloc first half", srcs, dsts, devs[: n // 2], "both")
        run("default-alloc second half", srcs, dsts, devs[n // 2:], "both")
    del srcs, dsts
    # (b) buffers bound to the GPU's own NUMA node
    keep = []
    srcs, dsts = {}, {}
    stat = []
    for g in devs:
        s, k1, st1 = numa_buffer(NBYTES, nodes[g])
        d, k2, st2 = numa_buffer(NBYTES, nodes[g])
        srcs[g], dsts[g] = s, d
        keep += [k1, k2]
        stat.append((g, nodes[g], st1, st2))
    print("numa-bound buffers:", stat)
    for which in ("h2d", "d2h", "both"):
        run("numa-local all gpus", srcs, dsts, devs, which)
    # (c) deliberately remote: node of GPU g swapped with the other socket's
    if len(set(nodes)) > 1 and min(nodes) >= 0:
        other = {a: b for a, b in zip(sorted(set(nodes)), reversed(sorted(set(nodes))))}
        srcs2, dsts2 = {}, {}
        for g in devs:
            s, k1, _ = numa_buffer(NBYTES, other[nodes[g]])
            d, k2, _ = numa_buffer(NBYTES, other[nodes[g]])
            srcs2[g], dsts2[g] = s, d
            keep += [k1, k2]
        run("numa-remote all gpus", srcs2, dsts2, devs, "both")


if __name__ == "__main__":
    main()
