#!/usr/bin/env python
"""Probe: what the box can feed its GPUs from page-locked host memory when 1 / 2 / 4 / 8 of them copy at the same time.

The end-to-end number of bench.py (cm_submit_clouds_pinned + cm_merge_frame_async + cm_wait_frame) is bound by these
copies: one process per GPU, each streaming its frames from its own pinned arena. This script measures the raw ceiling of
exactly that pattern -- one process per GPU, back-to-back 32 MB cudaMemcpyAsync from pinned memory, all processes released
by a barrier -- for several ways of placing the arena:
   default      cudaHostAlloc as the process finds it (first touch by the allocating thread)
   numa_local   the process is pinned to the CPUs of the GPU's NUMA node before it allocates and touches the arena
   wc           cudaHostAllocWriteCombined
   interleave   pages spread over all NUMA nodes (numactl-style, through mbind via libnuma when present)
and records where everything sits (GPU -> PCI bus id -> NUMA node, CPU list per node). One JSON object on stdout.

usage: python scripts/h2d_probe.py [--gpus 1,2,4,8] [--mb 32] [--seconds 1.0]
"""
import argparse
import ctypes
import json
import multiprocessing as mp
import os
import subprocess
import sys
import time


def gpu_numa(pci_bus_id: str):
    """NUMA node of a GPU from sysfs (-1: the platform does not say)."""
    bdf = pci_bus_id.lower()
    if len(bdf.split(":")[0]) == 8:   # 00000000:1B:00.0 -> 0000:1b:00.0
        bdf = bdf[4:]
    try:
        return int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
    except Exception:
        return -1


def node_cpus(node: int):
    try:
        txt = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip()
    except Exception:
        return []
    cpus = []
    for part in txt.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus += list(range(int(a), int(b) + 1))
        elif part:
            cpus.append(int(part))
    return cpus


def worker(rank, n_active, mode, mb, seconds, barrier, out_q, bus):
    try:
        import torch
        from cuda import cudart
        torch.cuda.set_device(rank)
        node = gpu_numa(bus)
        info = {"gpu": rank, "pci": bus, "numa_node": node, "mode": mode}
        if mode == "numa_local" and node >= 0:
            cpus = node_cpus(node)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["pinned_to_cpus"] = "%d..%d (%d)" % (cpus[0], cpus[-1], len(cpus))
        nbytes = mb << 20
        n_buf = 4
        flags = cudart.cudaHostAllocWriteCombined if mode == "wc" else cudart.cudaHostAllocDefault
        hosts = []
        for _ in range(n_buf):
            err, p = cudart.cudaHostAlloc(nbytes, flags)
            assert err == cudart.cudaError_t.cudaSuccess, err
            ctypes.memset(p, 1, nbytes)   # first touch by this (possibly pinned) thread
            hosts.append(p)
        dev = torch.empty(n_buf * nbytes, dtype=torch.uint8, device="cuda")
        stream = torch.cuda.Stream()
        kind = cudart.cudaMemcpyKind.cudaMemcpyHostToDevice

        def burst(count):
            for i in range(count):
                cudart.cudaMemcpyAsync(dev.data_ptr() + (i % n_buf) * nbytes, hosts[i % n_buf], nbytes, kind, stream.cuda_stream)
        burst(8)
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        count = max(8, int(seconds * 50e9 / nbytes))
        barrier.wait()
        with torch.cuda.stream(stream):
            e0.record(stream)
            burst(count)
            e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1)
        info.update(gbs=round(count * nbytes / (ms * 1e-3) / 1e9, 2), copies=count, copy_mb=mb)
        barrier.wait()
        for p in hosts:
            cudart.cudaFreeHost(p)
        out_q.put(info)
    except Exception as e:  # noqa: BLE001
        out_q.put({"gpu": rank, "error": repr(e)})
        try:
            barrier.abort()
        except Exception:
            pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1,2,4,8")
    ap.add_argument("--mb", type=int, default=32)
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--modes", default="default,numa_local,wc")
    a = ap.parse_args()
    import torch
    have = torch.cuda.device_count()
    res = {"what": "concurrent pinned H2D ceilings, one process per GPU, %d MB cudaMemcpyAsync back to back" % a.mb,
           "visible_gpus": have, "cpus": os.cpu_count(), "runs": []}
    try:
        res["topology"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout
        res["lscpu_numa"] = [l for l in subprocess.run(["lscpu"], capture_output=True, text=True, timeout=30).stdout.splitlines()
                             if "NUMA" in l or "Model name" in l or "Socket" in l]
    except Exception as e:  # noqa: BLE001
        res["topology_error"] = repr(e)
    try:
        res["meminfo_total_gb"] = round(int(open("/proc/meminfo").readline().split()[1]) / 1e6, 1)
    except Exception:
        pass
    try:
        buses = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True,
                               timeout=30).stdout.split()
    except Exception:
        buses = []
    buses += ["?"] * (have - len(buses))
    res["gpu_numa_nodes"] = [gpu_numa(b) for b in buses[:have]]
    ctx = mp.get_context("spawn")
    for mode in a.modes.split(","):
        for n in [int(x) for x in a.gpus.split(",")]:
            if n > have:
                continue
            barrier = ctx.Barrier(n)
            q = ctx.Queue()
            procs = [ctx.Process(target=worker, args=(r, n, mode, a.mb, a.seconds, barrier, q, buses[r])) for r in range(n)]
            for p in procs:
                p.start()
            got = []
            t0 = time.time()
            while len(got) < n and time.time() - t0 < 180:
                try:
                    got.append(q.get(timeout=5))
                except Exception:
                    if not any(p.is_alive() for p in procs):
                        break
            for p in procs:
                p.join(timeout=30)
                if p.is_alive():
                    p.kill()
            got.sort(key=lambda d: d.get("gpu", 0))
            rates = [g["gbs"] for g in got if "gbs" in g]
            res["runs"].append({"mode": mode, "gpus_copying": n, "per_gpu_gbs": rates, "aggregate_gbs": round(sum(rates), 1),
                                "min_gbs": min(rates) if rates else None, "detail": got})
            print("[h2d] %-10s %d GPUs: aggregate %.1f GB/s, per GPU %s" % (mode, n, sum(rates), rates), file=sys.stderr)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
