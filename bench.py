#!/usr/bin/env python
"""bench.py -- merged + voxel-filtered Mpoints/s of the merge hot path (BASELINE.json metric) on N B200s.

A step = one pass of the whole hot path (transform -> concat -> crop -> VoxelGrid) over one batch of F frames of the
named workload (default: BASELINE config 3, 8 sensors x 256k points per frame, box crop, VoxelGrid 0.05 m -- the config
BASELINE.json names for 1/2/4/8 GPUs; frames are sharded over the ranks with no collective).
  value    device-resident throughput: inputs already in HBM, CUDA-event timed, max over ranks
  e2e      the same metric through the host C-ABI path (cm_submit_cloud_pinned / cm_merge_frame_async / cm_wait_frame)
           from pinned HOST buffers, H2D and D2H inside the timed region
  roofline the dominant kernel's algorithmic bytes / its CUDA-event duration against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (port of the PCL 1.8.1 path; the reference's own PCL build is not installable here)
--impl reference times that CPU path alone, on all host cores, for the driver's own ratio.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "merged+voxel-filtered Mpoints/s"
UNIT = "Mpoints/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json config; cfg3 (8 x 256k points per frame, frame-sharded) is the 1/2/4/8-GPU config")
    ap.add_argument("--frames", type=int, default=0, help="frames per step (0 = enough to exceed L2 several times)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-path measurement (0 = min(steps, 3))")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-latency", action="store_true", help="skip the single-frame p50/p99 latency measurement")
    ap.add_argument("--sustained-seconds", type=float, default=2.0,
                    help="length of the extra back-to-back leg that shows the number holds at steady-state clocks (0 = skip)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe). The timed region of this
    workload lasts milliseconds, so the primary source is an NVML polling thread (one sample per ~2 ms); nvidia-smi -lms
    (one sample per 100 ms at best) runs beside it as the recipe's own view and is merged in."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, uuid: str = ""):
        self.gpu = gpu_index
        self.uuid = uuid
        self.rows = []
        self.proc = None
        self.nv = []        # (sm_mhz, reasons bitmask)
        self.nv_max = None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None
        self._nvml_open()

    def _nvml_open(self):
        """NVML handle of this rank's GPU, opened before the timed region (nvmlInit takes milliseconds)."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            if self.uuid:
                for cand in (self.uuid, "GPU-" + self.uuid):
                    try:
                        h = nv.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.nv_max = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self._nvml = (nv, h)
        except Exception as e:  # keep going on nvidia-smi alone, but say why
            self._nvml = None
            print("[bench] NVML sampling unavailable: %r" % (e,), file=sys.stderr)

    def _nvml_loop(self):
        if not self._nvml:
            return
        nv, h = self._nvml

        def get_reasons(hh):
            for name in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
                fn = getattr(nv, name, None)
                if fn is not None:
                    try:
                        return int(fn(hh))
                    except Exception:
                        continue
            return 0
        try:
            while not self._stop.is_set():
                self.nv.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), get_reasons(h)))
                time.sleep(0.001)
        except Exception as e:
            print("[bench] NVML sampling stopped: %r" % (e,), file=sys.stderr)

    def start(self):
        self._thr = threading.Thread(target=self._nvml_loop, daemon=True)
        self._thr.start()
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1.0)
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        # NVML bit masks: 0x8 hw_slowdown, 0x40 hw_thermal_slowdown, 0x20 sw_thermal_slowdown, 0x4 sw_power_cap
        for clk, bits in self.nv:
            sm.append(clk)
            for name, bit in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)):
                if bits & bit:
                    reasons.add(name)
        if self.nv_max:
            mx.append(self.nv_max)
        n_nvml = len(sm)
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvml x%d + nvidia-smi x%d" % (n_nvml, len(sm) - n_nvml)}


def dram_traffic(kernel: str, workload: str, frames: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed `ncu --set full` capture of
    this same workload (profiles/traffic.json, written by scripts/ncu_traffic.py); None when no capture matches."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        e = d.get("%s/%d" % (workload, frames), {}).get(kernel)
        return int(e["bytes_per_launch"]) if e else None
    except Exception:
        return None


def workload_spec(name: str, frames: int):
    from cloud_merger_b200 import synth
    c = dict(synth.CONFIGS[name])
    n = c["rings"] * c["azimuth"]
    if frames <= 0:
        # >= 4x the 126 MB L2 of raw input per step, so every step streams from HBM
        frames = max(1, int(np.ceil(4 * 126e6 / (c["sensors"] * n * 16))))
        frames = min(frames, 64)
    c.update(points_per_sensor=n, frames=frames)
    return c


def make_host_frames(spec, name: str, rank: int, pinned: bool):
    """F distinct synthetic frames (new noise per frame), packed 16-byte xyzi records, one buffer per (frame, sensor)."""
    from cloud_merger_b200 import host_alloc, synth
    S, n, F = spec["sensors"], spec["points_per_sensor"], spec["frames"]
    seed = 1000 * int(name[-1])
    nbytes = n * 16
    if pinned:
        arena, addr = host_alloc(F * S * nbytes)
    else:
        arena, addr = np.empty(F * S * nbytes, np.uint8), 0
    bufs = []
    for f in range(F):
        row = []
        for s in range(S):
            off = (f * S + s) * nbytes
            view = arena[off:off + nbytes].view(np.float32).reshape(n, 4)
            view[:] = synth.lidar_cloud(seed, s, rank * 100000 + f, spec["rings"], spec["azimuth"])
            row.append((view, (addr + off) if pinned else view.ctypes.data))
        bufs.append(row)
    return arena, bufs


def cpu_port(spec, host_frames, seconds: float, threads_frames: int, sensor_threads: int):
    """Times the CPU oracle (port of the reference's PCL path) on a bounded sample of the same frames."""
    from concurrent.futures import ThreadPoolExecutor

    from cloud_merger_b200 import synth
    from oracle import cm_oracle_py as oracle
    S = spec["sensors"]
    mats = [synth.extrinsic(s, S)[:3].reshape(-1) for s in range(S)]

    def one(f):
        cds = [dict(data=host_frames[f][s][0].view(np.uint8).reshape(-1), n_points=spec["points_per_sensor"], point_step=16,
                    off_x=0, off_y=4, off_z=8, off_i=12, is_dense=1, m=mats[s]) for s in range(S)]
        r = oracle.merge_frame(cds, spec["passes"], [spec["leaf"]] * 3, spec["min_points"], True, True,
                               threads=sensor_threads, want_outputs=False)
        return r["n_voxels"]

    F = len(host_frames)
    one(0)  # warm-up (page faults, library load)
    done, t0 = 0, time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads_frames) as ex:
        while True:
            batch = [(done + i) % F for i in range(threads_frames)]
            list(ex.map(one, batch))
            done += len(batch)
            if time.perf_counter() - t0 >= seconds:
                break
    dt = time.perf_counter() - t0
    pts = done * S * spec["points_per_sensor"]
    return pts / dt / 1e6, done, dt


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path. PCL/ROS cannot be installed here
    (SURVEY.md section 8c), so this is the oracle port, run frame-parallel on every host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    spec = workload_spec(args.workload, args.frames)
    cores = os.cpu_count() or 1
    sample_frames = min(spec["frames"], max(4, cores))
    spec_small = dict(spec, frames=sample_frames)
    _, frames = make_host_frames(spec_small, args.workload, 0, pinned=False)
    from concurrent.futures import ThreadPoolExecutor

    from cloud_merger_b200 import synth
    from oracle import cm_oracle_py as oracle
    S = spec["sensors"]
    mats = [synth.extrinsic(s, S)[:3].reshape(-1) for s in range(S)]

    def one(f):
        cds = [dict(data=frames[f][s][0].view(np.uint8).reshape(-1), n_points=spec["points_per_sensor"], point_step=16,
                    off_x=0, off_y=4, off_z=8, off_i=12, is_dense=1, m=mats[s]) for s in range(S)]
        return oracle.merge_frame(cds, spec["passes"], [spec["leaf"]] * 3, spec["min_points"], True, True, threads=1,
                                  want_outputs=False)["n_voxels"]

    workers = min(cores, sample_frames)
    with ThreadPoolExecutor(max_workers=workers) as ex:
        def step():
            list(ex.map(one, range(sample_frames)))
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
    pts_step = sample_frames * S * spec["points_per_sensor"]
    value = pts_step * args.steps / dt / 1e6
    sample = "%d frames/step of %s, frame-parallel on %d threads (oracle port of the PCL 1.8.1 path)" % (
        sample_frames, args.workload, workers)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, spec["what"]), "frames_per_step": sample_frames,
                   "points_per_frame": S * spec["points_per_sensor"], "leaf_m": spec["leaf"],
                   "min_points": spec["min_points"], "crop": spec["passes"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


_REAL_STDOUT = None


def guard_stdout():
    """The contract is ONE JSON line on stdout. Libraries print there too (NCCL's version banner at NCCL_DEBUG=WARN, for
    one), so file descriptor 1 is pointed at stderr for the whole run and the line is written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse_args()
    guard_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from cloud_merger_b200 import CloudMerger, make_layout, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    spec = workload_spec(args.workload, args.frames)
    S, n, F = spec["sensors"], spec["points_per_sensor"], spec["frames"]
    pts_step = F * S * n
    pinned_arena, host_frames = make_host_frames(spec, args.workload, rank, pinned=True)

    cm = CloudMerger(device=local_rank, max_sensors=S, max_points_per_sensor=n, max_point_step=16, frames_in_flight=4,
                     max_batch_points=pts_step, max_batch_frames=F)
    for s in range(S):
        cm.set_extrinsic(s, synth.extrinsic(s, S))
    cm.set_crop(spec["passes"])
    cm.set_voxel(spec["leaf"], spec["min_points"], True)
    layout = make_layout()

    # ---- device-resident inputs -------------------------------------------------------------------------------------
    dev = torch.empty(pts_step * 16, dtype=torch.uint8, device="cuda")
    dev.copy_(torch.from_numpy(pinned_arena), non_blocking=False)
    items = []
    for f in range(F):
        for s in range(S):
            items.append((dev.data_ptr() + (f * S + s) * n * 16, n, layout, s, f))
    segs = cm.make_segments(items)
    stream = torch.cuda.current_stream().cuda_stream
    cm.set_profiling(True)

    stage_names = ["transform_crop", "grid", "key_hist", "sort", "centroid"]
    stage_acc = {k: 0.0 for k in stage_names}
    for _ in range(max(args.warmup, 3)):
        cm.run_batch(segs, stream=stream)
        cm.sync()
    st = cm.stats()
    if st.device_error:
        raise SystemExit("device error %d" % st.device_error)

    try:
        gpu_uuid = str(torch.cuda.get_device_properties(torch.cuda.current_device()).uuid)
    except Exception:
        gpu_uuid = ""
    sampler = ClockSampler(local_rank, gpu_uuid)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    # The K steps are enqueued back to back, as a consumer of device-resident batches would, and the host reads the
    # report (survivor / voxel counts, stage events) once at the end: the stage times are those of the last timed step.
    ev0.record()
    for _ in range(args.steps):
        cm.run_batch(segs, stream=stream)
    ev1.record()
    cm.sync()
    launches = cm.launch_count() * args.steps
    for k in stage_names:
        stage_acc[k] = cm.stage_ms(k) * args.steps
    barrier()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = pts_step * world * args.steps / (ms_total / 1e3) / 1e6
    st = cm.stats()
    M, V, P, kb = int(st.survivors), int(st.voxels_out), int(st.sort_passes), int(st.key_bytes)

    # ---- sustained leg: the same step back to back for >= --sustained-seconds (the K-step region above lasts milliseconds,
    # too short for the GPU to reach its steady-state power / clocks) ---------------------------------------------------------
    sustained = None
    if args.sustained_seconds > 0:
        cm.set_profiling(False)
        n_sus = max(args.steps, int(np.ceil(args.sustained_seconds * 1e3 / ms_step)))
        sampler2 = ClockSampler(local_rank, gpu_uuid)
        barrier()
        sampler2.start()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for i in range(n_sus):
            cm.run_batch(segs, stream=stream)
        s1.record()
        cm.sync()
        barrier()
        clocks2 = sampler2.stop()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        sus_ms = float(ts.item())
        sustained = {"steps": n_sus, "gpu_launches": int(cm.launch_count() * n_sus), "seconds": sus_ms / 1e3, "ms_per_step": sus_ms / n_sus,
                     "value": pts_step * world * n_sus / (sus_ms / 1e3) / 1e6, "unit": UNIT, "clocks": clocks2}
        cm.set_profiling(True)

    # ---- roofline: algorithmic bytes per launch / CUDA-event duration of that launch ----------------------------------
    peak, peak_src = peaks()
    stage_ms = {k: v / args.steps for k, v in stage_acc.items()}
    algo = {
        "transform_crop": pts_step * 16 + M * 20,
        "key_hist": M * (16 + kb + 4),
        "sort": M * (2 * (kb + 4) * P),
        "centroid": M * (kb + 4 + 16) + V * 28,
    }
    stages = {}
    for k in ("transform_crop", "key_hist", "sort", "centroid"):
        ms = stage_ms[k]
        gbs = algo[k] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        stages[k] = {"ms": round(ms, 4), "algorithmic_bytes": int(algo[k]), "achieved_gbs": round(gbs, 1),
                     "frac": round(gbs / peak, 4), "launches": P if k == "sort" else 1}
    vg_ms = stage_ms["grid"] + stage_ms["key_hist"] + stage_ms["sort"] + stage_ms["centroid"]
    b_vg = M * (16 + 2 * kb + 2 * P * (kb + 4) + (kb + 4) + 16) + 20 * V   # SURVEY 8d named-pipeline traffic
    b_io = 16 * M + 20 * V
    stages["voxelgrid_end_to_end"] = {
        "ms": round(vg_ms, 4), "B_vg_bytes": int(b_vg), "achieved_gbs": round(b_vg / (vg_ms * 1e-3) / 1e9, 1) if vg_ms > 0 else 0,
        "frac": round(b_vg / (vg_ms * 1e-3) / 1e9 / peak, 4) if vg_ms > 0 else 0,
        "B_io_bytes": int(b_io), "frac_io": round(b_io / (vg_ms * 1e-3) / 1e9 / peak, 4) if vg_ms > 0 else 0,
        "key_bytes": kb, "passes": P}
    dom = max(("transform_crop", "key_hist", "sort", "centroid"), key=lambda k: stage_ms[k])
    dom_launches = P if dom == "sort" else 1
    dom_ms_launch = stage_ms[dom] / dom_launches
    dom_bytes_launch = algo[dom] / dom_launches
    achieved = dom_bytes_launch / (dom_ms_launch * 1e-3) / 1e9 if dom_ms_launch > 0 else 0.0
    roof_kernel = {"transform_crop": "k_transform_crop", "key_hist": "k_voxel_key_hist", "sort": "k_onesweep_pass",
                   "centroid": "k_voxel_centroid"}[dom]
    roofline = {"bound": "hbm", "kernel": roof_kernel,
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": dram_traffic(roof_kernel, args.workload, F), "peak_source": peak_src, "launch_ms": round(dom_ms_launch, 4),
                "algorithmic_bytes_per_launch": int(dom_bytes_launch), "share_of_step": round(stage_ms[dom] / ms_step, 3)}

    # ---- end to end through the host C-ABI path, pinned host buffers, H2D + D2H inside the timed region -----------------
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or min(args.steps, 3)
        cm.set_profiling(False)  # the stage events of the resident run are not wanted between the kernels of this path
        cap = S * n
        h2d = d2h = 0

        ring = [cm.make_frame_buffers(cap, want_survivors=False, pinned=True) for _ in range(4)]

        # one call per frame hands over all S page-locked sensor clouds (adjacent in host memory: one 8 MB PCIe copy)
        submit = [cm.prepared_frame_submit(list(range(S)), [host_frames[f][s][1] for s in range(S)], [n] * S,
                                           [layout] * S, stamp=f) for f in range(F)]

        def host_step(count_bytes: bool):
            nonlocal h2d, d2h
            pending = []
            for f in range(F):
                submit[f]()
                pending.append((cm.merge_frame_async(), ring[f % 4][0]))
                if len(pending) >= 3:
                    t, o = pending.pop(0)
                    cm.wait_frame_into(t, o)
                    if count_bytes:
                        d2h += o.n_voxels * 28
            while pending:
                t, o = pending.pop(0)
                cm.wait_frame_into(t, o)
                if count_bytes:
                    d2h += o.n_voxels * 28
            if count_bytes:
                h2d += F * S * n * 16

        host_step(False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step(True)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": pts_step * world * e2e_steps / dt / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d / e2e_steps), "d2h_bytes_per_step": int(d2h / e2e_steps),
               "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
               "api": "cm_submit_clouds_pinned + cm_merge_frame_async + cm_wait_frame, 3 frames in flight"}

    # ---- single-frame latency (resident inputs) -----------------------------------------------------------------------
    latency = None
    if not args.no_latency and rank == 0:
        one = cm.make_segments(items[:S])
        lat = []
        for i in range(220):
            cm.run_batch(one, stream=stream)
            cm.sync()
            if i >= 20:
                lat.append(cm.stats().gpu_ms)
        lat.sort()
        latency = {"p50_ms": lat[len(lat) // 2], "p99_ms": lat[int(len(lat) * 0.99)], "frames": len(lat),
                   "what": "device time of one resident frame, first to last kernel"}
        if not args.no_e2e:
            # the same through the host path, nothing else in flight: page-locked clouds in -> voxels in host memory
            hl = []
            for i in range(120):
                f = i % F
                t0 = time.perf_counter()
                submit[f]()
                cm.wait_frame_into(cm.merge_frame_async(), ring[0][0])
                if i >= 20:
                    hl.append((time.perf_counter() - t0) * 1e3)
            hl.sort()
            latency.update({"host_p50_ms": hl[len(hl) // 2], "host_p99_ms": hl[int(len(hl) * 0.99)],
                            "host_what": "wall time of one frame through cm_submit_clouds_pinned + merge + wait (8 MB in, voxels out)"})

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sensor_threads = min(S, 6)
        v, frames_done, secs = cpu_port(spec, host_frames, args.cpu_seconds, 1, sensor_threads)
        cpu = {"value": v, "unit": UNIT, "cores": sensor_threads, "kind": "port",
               "sample": "%d frames of %s in %.1f s; per-sensor transform+crop on %d threads, concat+VoxelGrid on 1 "
                         "(mirrors AsyncSpinner(6) + main loop); oracle port, PCL itself is not installable here" % (
                             frames_done, args.workload, secs, sensor_threads)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.workload, spec["what"]), "frames_per_step": F,
                       "points_per_frame": S * n, "points_per_step_per_gpu": pts_step, "leaf_m": spec["leaf"],
                       "min_points": spec["min_points"], "crop": spec["passes"], "survivors_per_step": M,
                       "voxels_per_step": V, "key_bytes": kb, "sort_passes": P, "key_bits": int(st.key_bits),
                       "sharding": "frames round-robin over ranks, no collective",
                       "timed_region": "K steps enqueued back to back on one stream, one report read at the end; stage times = last timed step",
                       "l2": "inputs larger than L2 (%.0f MB raw input per step per GPU)" % (pts_step * 16 / 1e6)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "stages": stages,
            "cpu_baseline": cpu, "sustained": sustained,
        }
        if latency:
            line["latency"] = latency
        emit(line)
    cm.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
