#!/usr/bin/env python
"""bench.py -- headline measurement of the hot path (BASELINE.json metric: Mrays/s closest-hit (primary + diffuse) and BVH
build Mtris/s) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2]

One STEP = one frame of the workload through the C ABI: primary rays of this rank's camera (bmTraceCamera, reference
semantics), one diffuse bounce generation (bmBounceRaysCamera) and its trace (bmTraceRaysIndirect: the bounce count stays on the
device); at N > 1 the ranks render interleaved rows of an N-times taller image with a replicated tree (weak scaling) and all-gather
the framebuffer (NCCL) under the bounce trace.  Multi-sample workloads (c3, c4): per sample bmCameraRays (jitter) -> bmTraceRaysTiled ->
bmBounceRays -> bmTraceRaysIndirect -> bmProgressiveAccumulate, then bmProgressiveResolve.
`value` = rays traced by all ranks / device time of the step with every input resident in HBM, each step bracketed by a synchronise;
`e2e` = the same step fed from pinned HOST buffers (camera table H2D each step, framebuffer + bounce count + bounce hits D2H each
step), frames pipelined three deep, the K steps timed as one bracket.  The BVH build (soup resident -> tree traceable) is timed in
the same run and reported under "build".  L2 is flushed between timed steps (inside the compute stream in the e2e leg).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
METRIC = "Mrays/s closest-hit (primary + diffuse)"
FALLBACK_HBM_GBS = 6650.0      # /opt/skills/guides/B200_PROFILING.md fallback, used only if MEASURED_PEAKS.json is absent
NODE_BYTES, TRI_BYTES, CLUSTER_BYTES, LEAF_BYTES = 96, 48, 32, 48     # bytes a visit reads: node = 6 float4 of its 128-byte record
# Primary rays are traced in the reference's semantics (BM_TRACE_REFERENCE = 0, the parity mode).  Diffuse bounce rays are traced as closest
# hits with t >= 0 (BM_TRACE_CLOSEST = 1): a bounce ray starts 1e-4 above its source triangle, and the reference's any-sign acceptance would
# "hit" that triangle's plane at t < 0 in the first leaf it enters (99 % of bounce rays; tests/test_gpu_trace.py pins this).
BOUNCE_MODE = 1


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).  NVML is polled from a thread every
    2 ms (an nvidia-smi loop cannot place a sample inside a 20 ms region, least of all with eight GPUs); nvidia-smi is the fallback."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index; self.proc = None; self.path = None
        self.thread = None; self.stop_flag = False; self.sm = []; self.mx = None; self.bits = 0; self.nvml = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis and all(x.strip().isdigit() for x in vis.split(",")):
            return int(vis.split(",")[self.index])
        return self.index

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.bits |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml
            import threading
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None; self.thread = None
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False); self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self._physical_index()), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            n = self.nvml
            try:                                          # one more sample from this thread, the instant the timed region ends
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                self.bits |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:
                pass
            self.stop_flag = True
            self.thread.join(timeout=2)
            names = (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"), ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                     ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"), ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"))
            out["reasons"] = sorted(k for k, a in names if self.bits & int(getattr(n, a, 0)))
            if self.sm:
                out.update(sm_mhz=statistics.median(self.sm), sm_max_mhz=self.mx, samples=len(self.sm))
            out["source"] = "nvml"
            return out
        if not self.proc:
            out["reasons"] = ["nvidia-smi unavailable"]; return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        out["source"] = "nvidia-smi"
        return out


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ---------------------------------------------------------------------------------------------------------------------------
def bind_near_gpu(index):
    """Pin this process to the CPU cores of the GPU's NUMA node (NVML's ideal affinity) BEFORE any pinned host buffer is allocated:
    first-touch then places the staging buffers on the socket the GPU hangs off, so the H2D / D2H copies of N ranks do not cross the
    inter-socket link.  Best effort; returns a short description for the bench line."""
    if os.environ.get("BENCH_NO_AFFINITY"):
        return "off"
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%d cpus near gpu %d" % (len(cpus), phys)
    except Exception as ex:
        return "unavailable (%s)" % type(ex).__name__
    return "unavailable"


def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (the product has no CPU path)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa = bind_near_gpu(local)
    # libraries (NCCL's version banner) must not write to stdout: the driver reads ONE JSON line there
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from raytracercuda_b200 import build
    if local == 0:
        build.build_library()
    if world > 1:
        dist.barrier()
    from raytracercuda_b200 import _lib, host, workloads
    from raytracercuda_b200.tree import Tree, make_mesh_table
    L = _lib.load()
    wl = workloads.get(args.workload)
    W, H = wl.width, wl.height
    n = W * H
    eye = wl.eye; orient = np.eye(3, dtype=np.float32)
    frustum = workloads.interleaved_rows_frustum(wl.frustum, H, rank, world)

    # --- scene + camera resident on the device -------------------------------------------------------------------------
    cam = host.ICamera.create()
    assert cam.setInitialRays(W, H, *frustum) == 0
    rays_pinned = torch.from_numpy(cam.initialRays()).pin_memory()
    rays_dev = rays_pinned.to(dev)
    Pd = torch.from_numpy(wl.positions).to(dev); Id = torch.from_numpy(wl.indices.view(np.int32)).to(dev).reshape(-1); Nd = torch.from_numpy(wl.normals).to(dev)
    table = make_mesh_table([(Pd, Id, Nd)], dev)
    stream = torch.cuda.current_stream(dev)
    # the tree is built once (rank 0) and replicated to the peers over NCCL (index-based arrays, bmTreeExport / bmTreeImport)
    built = Tree(dev) if rank == 0 else None

    def build_once():
        built.reset(); built.add_mesh(Pd, Id); built.build()

    if rank == 0:
        build_once()
    from raytracercuda_b200.dist import replicate_tree
    tree = replicate_tree(built, dev, 0)
    stats_t = torch.zeros(8, dtype=torch.int64, device=dev)
    if rank == 0:
        s0 = built.stats()
        stats_t.copy_(torch.tensor([s0[k] for k in ("triangles", "pairs", "leaves", "nodes", "depth", "max_leaf_faces", "overflow_leaves", "meshes")]))
    if world > 1:
        dist.broadcast(stats_t, 0)
    st = dict(zip(("triangles", "pairs", "leaves", "nodes", "depth", "max_leaf_faces", "overflow_leaves", "meshes"), [int(x) for x in stats_t.cpu()]))
    hits1 = torch.empty((n, 4), dtype=torch.float32, device=dev); pixels = torch.empty(n, dtype=torch.int32, device=dev)
    rays1 = torch.empty((n, 6), dtype=torch.float32, device=dev); rays2 = torch.empty((n, 6), dtype=torch.float32, device=dev)
    src2 = torch.empty(n, dtype=torch.int32, device=dev); cnt2 = torch.zeros(1, dtype=torch.int32, device=dev)
    hits2 = torch.empty((n, 4), dtype=torch.float32, device=dev)
    gathered = torch.empty((world, n), dtype=torch.int32, device=dev) if world > 1 else None
    comm = torch.cuda.Stream(dev) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    pix_host = torch.empty(n, dtype=torch.int32).pin_memory(); hits2_host = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    cnt_host = torch.zeros(1, dtype=torch.int32).pin_memory()
    ev = lambda: torch.cuda.Event(enable_timing=True)

    SPP = wl.samples
    jit = (abs(wl.frustum[1] - wl.frustum[0]) / W / wl.frustum[4], abs(frustum[3] - frustum[2]) / H / wl.frustum[4])
    acc = torch.zeros(n, dtype=torch.float32, device=dev) if SPP > 1 else None
    cov = torch.zeros(n, dtype=torch.int32, device=dev) if SPP > 1 else None
    cnt_sum = torch.zeros(1, dtype=torch.int64, device=dev)

    def frame_spp(rays_in):
        """multi-sample frame (configs 3 and 4): per sample jittered camera rays -> trace -> bounce rays -> trace -> accumulate; resolve"""
        acc.zero_(); cov.zero_(); cnt_sum.zero_()
        sp = C.c_void_p(stream.cuda_stream)
        for sidx in range(SPP):
            # the sample's jitter is applied where the trace kernel forms the ray (no 24-byte ray record per sample is written or read);
            # the launch is a camera launch: 8x4 tiles under the temporal tile schedule
            _lib.check(L.bmTraceCameraJittered(tree.h, rays_in.data_ptr(), n, W, _f3(eye), _f9(orient), jit[0], jit[1], sidx, args.seed, 0, hits1.data_ptr(), None, None, None), "bmTraceCameraJittered")
            _lib.check(L.bmBounceRaysCameraJittered(tree.h, rays_in.data_ptr(), _f3(eye), _f9(orient), jit[0], jit[1], sidx, args.seed, hits1.data_ptr(), n, sidx, 1, args.seed,
                                                    rays2.data_ptr(), src2.data_ptr(), cnt2.data_ptr()), "bmBounceRaysCameraJittered")
            _lib.check(L.bmTraceRaysIndirect(tree.h, rays2.data_ptr(), cnt2.data_ptr(), n, BOUNCE_MODE, hits2.data_ptr(), None), "bmTraceRaysIndirect")
            _lib.check(L.bmProgressiveAccumulate(hits1.data_ptr(), n, hits2.data_ptr(), src2.data_ptr(), cnt2.data_ptr(), 0.25, acc.data_ptr(), cov.data_ptr(), sp), "accumulate")
            cnt_sum.add_(cnt2)
        _lib.check(L.bmProgressiveResolve(acc.data_ptr(), cov.data_ptr(), n, SPP, pixels.data_ptr(), sp), "resolve")
        cnt_host.copy_(cnt_sum.to(torch.int32), non_blocking=True)
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), pixels)
        return n * SPP

    # The frame is issued as NB row bands on NB streams (each through its own view of the tree): the traversal kernels are persistent
    # grids with a long tail (a few grazing rays), and the tail of one band is filled by the other bands' work.  NB = 1 is one launch.
    NB = args.bands if (SPP == 1 and H % args.bands == 0) else 1
    nbnd = n // NB
    band_streams = [torch.cuda.Stream(dev) for _ in range(NB)] if NB > 1 else [stream]
    band_trees = [tree.view(band_streams[k]) for k in range(NB)] if NB > 1 else [tree]
    cntb = torch.zeros(NB, dtype=torch.int32, device=dev); cntb_host = torch.zeros(NB, dtype=torch.int32).pin_memory()

    def frame(rays_in, marks=None, out=None):
        """one step on the device; returns the primary rays traced by this rank (bounce rays are added by bounce_count())"""
        if SPP > 1:
            if marks is not None:
                for m_ in marks: m_.record()
            return frame_spp(rays_in)
        o_hits1, o_pixels, o_rays2, o_src2, o_hits2, o_cnt, o_cnt_host = out if out is not None else (hits1, pixels, rays2, src2, hits2, cntb, cntb_host)
        if marks is not None: marks[0].record()
        ev0 = torch.cuda.Event(); ev0.record(stream)
        prim_done = []
        for k in range(NB):
            lo = k * nbnd; bs = band_streams[k]; bt = band_trees[k]
            if NB > 1: bs.wait_event(ev0)
            _lib.check(L.bmTraceCamera(bt.h, rays_in.data_ptr() + lo * 12, nbnd, W, _f3(eye), _f9(orient), 0, o_hits1.data_ptr() + lo * 16, o_pixels.data_ptr() + lo * 4,
                                       table.data_ptr(), None), "bmTraceCamera")
            if marks is not None and NB == 1: marks[1].record()
            if world > 1:
                e_ = torch.cuda.Event(); e_.record(bs); prim_done.append(e_)
            if wl.bounces:
                _lib.check(L.bmBounceRaysCamera(bt.h, rays_in.data_ptr() + lo * 12, _f3(eye), _f9(orient), o_hits1.data_ptr() + lo * 16, nbnd, 0, 1, args.seed,
                                                o_rays2.data_ptr() + lo * 24, o_src2.data_ptr() + lo * 4, o_cnt.data_ptr() + k * 4), "bmBounceRaysCamera")
                if marks is not None and NB == 1: marks[2].record()
                # the bounce count stays on the device: the trace kernel reads it there (no host round trip inside the step)
                _lib.check(L.bmTraceRaysIndirect(bt.h, o_rays2.data_ptr() + lo * 24, o_cnt.data_ptr() + k * 4, nbnd, BOUNCE_MODE, o_hits2.data_ptr() + lo * 16, None), "bmTraceRaysIndirect")
            if NB > 1:
                e_ = torch.cuda.Event(); e_.record(bs); stream.wait_event(e_)
        work = None
        if world > 1:
            # the framebuffer is final once the primary launches are done: gather it on a side stream while the bounce rays are traced
            with torch.cuda.stream(comm):
                for e_ in prim_done: comm.wait_event(e_)
                work = dist.all_gather_into_tensor(gathered.view(-1), o_pixels, async_op=True)
        if wl.bounces:
            o_cnt_host.copy_(o_cnt, non_blocking=True)
        if marks is not None: marks[3].record()
        if work is not None:
            work.wait()                    # the step ends when the gathered frame is complete
        return n

    def bounce_count():
        stream.synchronize()
        if SPP > 1:
            return int(cnt_host[0]) if wl.bounces else 0
        return int(cntb_host.sum()) if wl.bounces else 0

    import ctypes as C
    _f3 = lambda a: (C.c_float * 3)(*[float(x) for x in a])
    _f9 = lambda a: (C.c_float * 9)(*[float(x) for x in np.asarray(a, np.float32).reshape(-1)])

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # --- warm-up -------------------------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        rays_step = frame(rays_dev) + bounce_count()
    # per-ray traversal counters of the same traversal (instrumented launch, outside the timed region)
    cnt = torch.zeros(4, dtype=torch.int64, device=dev)
    _lib.check(L.bmTraceCamera(tree.h, rays_dev.data_ptr(), n, W, _f3(eye), _f9(orient), 0, hits1.data_ptr(), None, None, cnt.data_ptr()), "bmTraceCamera")
    torch.cuda.synchronize(dev)
    V1, K1, C1, LONG1 = [int(x) for x in cnt.cpu()]
    def whole_frame_bounce_list():
        """the frame's bounce rays as ONE compacted list in rays2 / src2 (the banded step leaves one list per band)"""
        _lib.check(L.bmBounceRaysCamera(tree.h, rays_dev.data_ptr(), _f3(eye), _f9(orient), hits1.data_ptr(), n, 0, 1, args.seed, rays2.data_ptr(),
                                        src2.data_ptr(), cnt2.data_ptr()), "bmBounceRaysCamera")
        torch.cuda.synchronize(dev)
        return int(cnt2.cpu()[0])
    if SPP == 1:
        nb = whole_frame_bounce_list() if wl.bounces else 0
        assert nb == (int(cntb_host.sum()) if wl.bounces else 0), "band lists and the whole-frame list disagree"
    else:
        nb = int(cnt_host[0]) if wl.bounces else 0
    V2 = K2 = C2 = 0
    nb_last = int(cnt2.cpu()[0]) if wl.bounces else 0      # rays2 holds the bounce rays of the LAST sample only
    if nb_last and SPP == 1:
        cnt.zero_()
        _lib.check(L.bmTraceRays(tree.h, rays2.data_ptr(), nb_last, BOUNCE_MODE, hits2.data_ptr(), cnt.data_ptr()), "bmTraceRays")
        torch.cuda.synchronize(dev)
        V2, K2, C2, LONG2 = [int(x) for x in cnt.cpu()]

    # --- timed: K steps, device resident, L2 flushed between steps --------------------------------------------------------
    sampler = ClockSampler(local)
    launches0 = L.bmLaunchCount()
    sync_all()
    sampler.start()
    step_ms, prim_ms, bounce_ms = [], [], []
    total_rays = 0
    for _ in range(args.steps):
        flush.zero_()
        marks = [ev() for _ in range(4)]; e0, e1 = ev(), ev()
        sync_all()
        e0.record()
        total_rays += frame(rays_dev, marks)
        e1.record()
        sync_all()
        total_rays += bounce_count()
        step_ms.append(e0.elapsed_time(e1)); prim_ms.append(marks[0].elapsed_time(marks[1]) if (NB == 1 or SPP > 1) else 0.0)
        bounce_ms.append(marks[2].elapsed_time(marks[3]) if (wl.bounces and (NB == 1 or SPP > 1)) else 0.0)
    clocks = sampler.stop()
    launches = L.bmLaunchCount() - launches0
    # max over ranks of the summed step time; rays summed over ranks
    t_sum = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev); r_sum = torch.tensor([float(total_rays)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_sum, op=dist.ReduceOp.MAX); dist.all_reduce(r_sum, op=dist.ReduceOp.SUM)
    ms_per_step = float(t_sum.item()) / args.steps
    value = float(r_sum.item()) / args.steps / (ms_per_step * 1e-3) / 1e6

    # --- timed: e2e through host buffers ---------------------------------------------------------------------------------
    # Every step: the camera table comes from pinned host memory (H2D), the framebuffer, the bounce count and the bounce hit records go
    # back to pinned host memory (D2H).  Frames are pipelined the way a renderer streams them: DEPTH buffer sets, three streams -- the
    # upload of frame i+1 and the download of frame i-1 run while frame i is traced.  The K steps are timed as one bracket (fill and
    # drain included); the latency of one un-pipelined frame is reported beside it.  The L2 flush runs inside the compute stream.
    DEPTH = 3
    s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    slots = []
    for _ in range(DEPTH if SPP == 1 else 0):
        slots.append({"rays": torch.empty_like(rays_dev), "hits1": torch.empty_like(hits1), "pixels": torch.empty_like(pixels), "rays2": torch.empty_like(rays2),
                      "src2": torch.empty_like(src2), "cnt": torch.zeros(NB, dtype=torch.int32, device=dev), "hits2": torch.empty_like(hits2),
                      "pix_host": torch.empty(n, dtype=torch.int32).pin_memory(), "hits2_host": torch.empty((n, 4), dtype=torch.float32).pin_memory(),
                      "cnt_host": torch.zeros(NB, dtype=torch.int32).pin_memory(), "cnt_scratch": torch.zeros(NB, dtype=torch.int32).pin_memory(),
                      "ev_comp": None, "ev_d2h": None, "ev_cnt": None})

    def e2e_upload(i):
        """camera table of frame i, host -> device (asynchronous; issued two frames ahead so that it never gates the kernels)"""
        S = slots[i % DEPTH]
        with torch.cuda.stream(s_h2d):
            if S["ev_comp"] is not None:
                s_h2d.wait_event(S["ev_comp"])              # the previous frame in this slot has consumed its ray table
            S["rays"].copy_(rays_pinned, non_blocking=True)
            S["ev_h2d"] = torch.cuda.Event(); S["ev_h2d"].record(s_h2d)

    e2e_graphs = None                                       # one CUDA graph per buffer set (filled in below): L2 flush + the frame's launches

    def frame_calls(S):
        """the frame's three C-ABI calls on buffer set S (single GPU, one band): what the per-slot CUDA graph captures"""
        _lib.check(L.bmTraceCamera(tree.h, S["rays"].data_ptr(), n, W, _f3(eye), _f9(orient), 0, S["hits1"].data_ptr(), S["pixels"].data_ptr(), table.data_ptr(), None), "bmTraceCamera")
        if wl.bounces:
            _lib.check(L.bmBounceRaysCamera(tree.h, S["rays"].data_ptr(), _f3(eye), _f9(orient), S["hits1"].data_ptr(), n, 0, 1, args.seed, S["rays2"].data_ptr(),
                                            S["src2"].data_ptr(), S["cnt"].data_ptr()), "bmBounceRaysCamera")
            _lib.check(L.bmTraceRaysIndirect(tree.h, S["rays2"].data_ptr(), S["cnt"].data_ptr(), n, BOUNCE_MODE, S["hits2"].data_ptr(), None), "bmTraceRaysIndirect")

    def e2e_compute(i):
        """kernels of frame i (asynchronous)"""
        S = slots[i % DEPTH]
        with torch.cuda.stream(stream):
            stream.wait_event(S["ev_h2d"])
            if S["ev_d2h"] is not None:
                stream.wait_event(S["ev_d2h"])              # ... and its results have left the device
            if e2e_graphs is not None:
                e2e_graphs[i % DEPTH].replay()
            else:
                flush.zero_()
                frame(S["rays"], None, out=(S["hits1"], S["pixels"], S["rays2"], S["src2"], S["hits2"], S["cnt"], S["cnt_scratch"]))
            S["ev_comp"] = torch.cuda.Event(); S["ev_comp"].record(stream)

    def e2e_download(i):
        """framebuffer + bounce count of frame i to the host (asynchronous)"""
        S = slots[i % DEPTH]
        with torch.cuda.stream(s_d2h):
            s_d2h.wait_event(S["ev_comp"])
            S["pix_host"].copy_(S["pixels"], non_blocking=True)
            S["cnt_host"].copy_(S["cnt"], non_blocking=True)
            S["ev_cnt"] = torch.cuda.Event(); S["ev_cnt"].record(s_d2h)

    def e2e_finish(i):
        """host side of frame i: wait for its bounce count, fetch exactly that many hit records; returns the rays of the frame"""
        S = slots[i % DEPTH]
        S["ev_cnt"].synchronize()
        total = n
        with torch.cuda.stream(s_d2h):
            for k in range(NB if wl.bounces else 0):
                c = int(S["cnt_host"][k]); lo = k * nbnd
                total += c
                if c:
                    S["hits2_host"][lo:lo + c].copy_(S["hits2"][lo:lo + c], non_blocking=True)
            S["ev_d2h"] = torch.cuda.Event(); S["ev_d2h"].record(s_d2h)
        return total

    def e2e_run(K):
        total = 0
        for i in range(min(2, K)):
            e2e_upload(i)
        for i in range(K):
            e2e_compute(i)
            if i + 2 < K:
                e2e_upload(i + 2)
            if i:
                total += e2e_finish(i - 1)
            e2e_download(i)
        total += e2e_finish(K - 1)
        stream.wait_event(slots[(K - 1) % DEPTH]["ev_d2h"])
        return total

    def frame_e2e_spp():
        e2e_rays_dev.copy_(rays_pinned, non_blocking=True)
        total = frame_spp(e2e_rays_dev)
        pix_host.copy_(pixels, non_blocking=True)
        stream.synchronize()
        return total + int(cnt_host[0])

    e2e_rays = 0
    if SPP == 1:
        e2e_run(DEPTH + 1)                                   # warm-up (allocations, all slots touched)
        if world == 1 and NB == 1 and not args.no_e2e_graph:
            # the launch sequence of a frame is fixed: captured once per buffer set, replayed as one graph launch per frame
            try:
                torch.cuda.synchronize(dev)
                graphs_ = []
                for S in slots:
                    g_ = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g_):
                        tree.use_stream(torch.cuda.current_stream(dev))
                        flush.zero_()
                        frame_calls(S)
                    graphs_.append(g_)
                tree.use_stream(stream)
                e2e_graphs = graphs_
                e2e_run(DEPTH + 1)
            except Exception as ex:                          # keep the run alive; the line says which path ran
                sys.stderr.write("e2e: CUDA graph capture failed (%r): direct launches\n" % (ex,))
                e2e_graphs = None
                tree.use_stream(stream)
        lat = []
        for _ in range(3):                                   # latency of one frame on its own
            e0, e1 = ev(), ev(); sync_all(); e0.record(); e2e_run(1); e1.record(); sync_all(); lat.append(e0.elapsed_time(e1))
        e0, e1 = ev(), ev()
        sync_all()
        e0.record(); e2e_rays = e2e_run(args.steps); e1.record()
        sync_all()
        e2e_total_ms = e0.elapsed_time(e1)
        e2e_latency_ms = statistics.mean(lat)
    else:
        e2e_rays_dev = torch.empty_like(rays_dev)
        for _ in range(2):
            frame_e2e_spp()
        e2e_total_ms = 0.0
        for _ in range(args.steps):
            flush.zero_()
            e0, e1 = ev(), ev()
            sync_all()
            e0.record(); e2e_rays += frame_e2e_spp(); e1.record()
            sync_all()
            e2e_total_ms += e0.elapsed_time(e1)
        e2e_latency_ms = e2e_total_ms / args.steps
    t2 = torch.tensor([e2e_total_ms], dtype=torch.float64, device=dev); r2 = torch.tensor([float(e2e_rays)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX); dist.all_reduce(r2, op=dist.ReduceOp.SUM)
    e2e_value = float(r2.item()) / (float(t2.item()) * 1e-3) / 1e6
    h2d = n * 12
    d2h = n * 4 + 4 * NB + (nb * 16 if (wl.bounces and SPP == 1) else 0)

    # --- timed: BVH build (soup resident -> tree traceable) ---------------------------------------------------------------
    build_ms = []
    if rank == 0:
        for i in range(2 + args.steps):
            flush.zero_()
            e0, e1 = ev(), ev()
            torch.cuda.synchronize(dev)
            e0.record(); build_once(); e1.record()
            torch.cuda.synchronize(dev)
            if i >= 2:
                build_ms.append(e0.elapsed_time(e1))
    rep_ms = 0.0
    if world > 1:
        # replication cost: export + NCCL broadcast of the arrays + import, max over ranks
        sync_all()
        reps = []
        for i in range(3):
            e0, e1 = ev(), ev()
            sync_all()
            e0.record(); _t = replicate_tree(built, dev, 0); e1.record()
            sync_all()
            reps.append(e0.elapsed_time(e1))
            if rank != 0:
                _t.close()
        rt = torch.tensor([min(reps)], dtype=torch.float64, device=dev)
        dist.all_reduce(rt, op=dist.ReduceOp.MAX)
        rep_ms = float(rt.item())
    bt = torch.tensor([statistics.mean(build_ms) if build_ms else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.broadcast(bt, 0)
    bms = float(bt.item())
    peak, peak_src = measured_peaks()
    build_bytes = 48 * st["triangles"] + 12 * st["pairs"] + 128 * st["nodes"] + 48 * st["pairs"] + 48 * st["leaves"] + 8 * st["pairs"]
    build = {"metric": "BVH build Mtris/s", "value": st["triangles"] / (bms * 1e-3) / 1e6, "unit": "Mtris/s", "ms": bms,
             "triangles": st["triangles"], "pairs": st["pairs"], "leaves": st["leaves"], "nodes": st["nodes"],
             "roofline": {"bound": "hbm", "achieved": build_bytes / (bms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": build_bytes / (bms * 1e-3) / 1e9 / peak, "traffic": None,
                          "algorithmic_bytes": build_bytes, "note": "compulsory bytes of the whole build: 48 B/tri soup read + 12 B/pair (8 B key + 4 B id) written + outputs (128 B/node, 48 B/leaf, 48 B/pair triangle record, 32 B per 4 pairs cluster box); sort passes excluded (SURVEY 8d); " + peak_src},
             "replication": ("built on rank 0, arrays replicated to %d peers by NCCL broadcast in %.3f ms (not included in ms)" % (world - 1, rep_ms)) if world > 1 else "single GPU",
             "replicate_ms": rep_ms}

    # --- roofline of the dominant kernel: the primary-ray trace launch ---------------------------------------------------
    pms = statistics.mean(prim_ms)
    if SPP > 1 or NB > 1:
        # multi-sample / banded frames interleave launches: time the dominant kernel (one camera launch of the frame's rays) on its own
        tt = []
        for _ in range(5):
            flush.zero_(); e0, e1 = ev(), ev(); torch.cuda.synchronize(dev); e0.record()
            _lib.check(L.bmTraceCamera(tree.h, rays_dev.data_ptr(), n, W, _f3(eye), _f9(orient), 0, hits1.data_ptr(), None, None, None), "bmTraceCamera")
            e1.record(); torch.cuda.synchronize(dev); tt.append(e0.elapsed_time(e1))
        pms = statistics.mean(tt)
    prim_bytes = n * (12 + 16 + 4) + NODE_BYTES * V1 + TRI_BYTES * K1 + CLUSTER_BYTES * C1
    roofline = {"bound": "hbm", "kernel": "trace_kernel<camera> (bmTraceCamera, reference semantics)", "achieved": prim_bytes / (pms * 1e-3) / 1e9,
                "peak": peak, "unit": "GB/s", "frac": prim_bytes / (pms * 1e-3) / 1e9 / peak, "traffic": None,
                "ms_per_launch": pms, "mrays_per_s": n / (pms * 1e-3) / 1e6,
                "bytes_per_ray": prim_bytes / n, "node_visits_per_ray": V1 / n, "tri_tests_per_ray": K1 / n, "cluster_boxes_per_ray": C1 / n, "longest_ray_steps": LONG1,
                "note": "algorithmic bytes = 12 B ray + 16 B hit + 4 B pixel + 96 B x node/leaf visits + 32 B x cluster boxes + 48 B x triangle tests (counters of the same traversal); "
                        + peak_src + "; the tree (~%.0f MB) is L2-resident on B200, so this kernel is latency/issue-bound, not HBM-bound (SURVEY 8d)" %
                        ((st["nodes"] * 128 + st["pairs"] * 56 + st["leaves"] * 48) / 1e6)}
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
        if tr:
            roofline["traffic"] = tr["dram_bytes_per_launch"]; roofline["traffic_source"] = tr["source"]
    except Exception:
        pass
    if wl.bounces and nb and SPP == 1:
        if NB > 1:
            bounce_ms = []
            for _ in range(5):
                flush.zero_(); e0, e1 = ev(), ev(); torch.cuda.synchronize(dev); e0.record()
                _lib.check(L.bmTraceRays(tree.h, rays2.data_ptr(), nb, BOUNCE_MODE, hits2.data_ptr(), None), "bmTraceRays")
                e1.record(); torch.cuda.synchronize(dev); bounce_ms.append(e0.elapsed_time(e1))
        bms2 = statistics.mean(bounce_ms)
        roofline["bounce_launch"] = {"ms_per_launch": bms2, "mrays_per_s": nb / (bms2 * 1e-3) / 1e6, "rays": nb,
                                     "node_visits_per_ray": V2 / nb, "tri_tests_per_ray": K2 / nb, "cluster_boxes_per_ray": C2 / nb,
                                     "achieved": (nb * (24 + 16) + NODE_BYTES * V2 + TRI_BYTES * K2 + CLUSTER_BYTES * C2) / (bms2 * 1e-3) / 1e9}

    line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "rays_per_step_per_gpu": rays_step, "primary_rays_per_gpu": n * SPP, "bounce_rays_per_gpu": nb, "samples_per_pixel": SPP,
                       "trace_mode": "primary rays: reference semantics (BM_TRACE_REFERENCE); bounce rays: closest hit, t >= 0 (BM_TRACE_CLOSEST)", "l2": "flushed between timed steps (256 MiB write)", "bands": NB, "cpu_affinity": numa,
                       "sharding": "interleaved image rows per rank, replicated tree, NCCL all-gather of the framebuffer overlapped with the bounce trace" if world > 1 else "single GPU",
                       "camera": {"eye": list(eye), "frustum": list(wl.frustum), "orient": "identity"}},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": float(t2.item()) / args.steps,
                    "single_frame_latency_ms": e2e_latency_ms,
                    "pipeline": ("%d frames in flight on 3 streams (H2D of the camera table | L2 flush + trace + bounce + trace%s | D2H of framebuffer, bounce count and "
                                 "bounce hit records sized by the device-side count); the K steps are one timed bracket, fill and drain included"
                                 % (DEPTH, ", replayed as one CUDA graph per frame" if e2e_graphs is not None else "")) if SPP == 1
                                else "one frame at a time: H2D of the camera table, the multi-sample frame, D2H of the framebuffer"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "build": build}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if SPP == 1 and wl.bounces:
            whole_frame_bounce_list()
        def gpu_frame_pixels(w_, h_):
            """the product's pixels for the workload's scene at w_ x h_ (same camera), for the comparison with the reference build"""
            c_ = host.ICamera.create(); assert c_.setInitialRays(w_, h_, *wl.frustum) == 0
            r_ = torch.from_numpy(c_.initialRays()).to(dev); p_ = torch.empty(w_ * h_, dtype=torch.int32, device=dev)
            _lib.check(L.bmTraceCamera(tree.h, r_.data_ptr(), w_ * h_, w_, _f3(eye), _f9(orient), 0, None, p_.data_ptr(), table.data_ptr(), None), "bmTraceCamera")
            torch.cuda.synchronize(dev)
            return p_.cpu().numpy().view(np.uint32)
        line["cpu_baseline"] = cpu_baseline(wl, rays_pinned.numpy(), rays2[:nb].cpu().numpy() if nb else None, src2[:nb].cpu().numpy() if nb else None,
                                            gpu_pixels_fn=gpu_frame_pixels)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.barrier()
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------------
def run_c5(args):
    """BASELINE config 5: counter-based triangle soup (default 100 M triangles) built once, then incoherent ray batches (origin uniform
    in the world, direction uniform on the sphere) traced in reference mode.  One step = one batch of --c5-mrays million rays per rank;
    ranks trace different batches against a replicated tree (weak scaling).  This is the HBM-bound case: the tree is far larger than L2."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    sys.stdout.flush(); saved_stdout = os.dup(1); os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from raytracercuda_b200 import build
    if local == 0:
        build.build_library()
    if world > 1:
        dist.barrier()
    from raytracercuda_b200 import _lib, synth
    from raytracercuda_b200.dist import replicate_tree
    from raytracercuda_b200.tree import Tree
    L = _lib.load()
    T = int(args.c5_mtris * 1e6); n = int(args.c5_mrays * 1e6)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    built, build_ms, st = None, [], None
    if rank == 0:
        Pd, Id = synth.soup_torch(T, dev)
        built = Tree(dev)
        for i in range(3):
            built.reset(); built.add_mesh(Pd, Id)
            e0, e1 = ev(), ev(); torch.cuda.synchronize(dev); e0.record(); built.build(); e1.record(); torch.cuda.synchronize(dev)
            build_ms.append(e0.elapsed_time(e1))
        st = built.stats()
    e0, e1 = ev(), ev(); torch.cuda.synchronize(dev)
    if world > 1: dist.barrier()
    e0.record(); tree = replicate_tree(built, dev, 0); e1.record(); torch.cuda.synchronize(dev)
    rep_ms = e0.elapsed_time(e1) if world > 1 else 0.0
    stt = torch.zeros(4, dtype=torch.int64, device=dev)
    if rank == 0:
        stt.copy_(torch.tensor([st["triangles"], st["pairs"], st["leaves"], st["nodes"]]))
    if world > 1:
        dist.broadcast(stt, 0)
    Tn, Pn, Ln, Nn = [int(x) for x in stt.cpu()]
    rays = synth.incoherent_rays_torch(n, dev, start=rank * n)
    hits = torch.empty((n, 4), dtype=torch.float32, device=dev)
    rays_pinned = rays.cpu().pin_memory(); hits_host = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream(dev)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(); torch.cuda.synchronize(dev)

    def step(src):
        _lib.check(L.bmTraceRays(tree.h, src.data_ptr(), n, 0, hits.data_ptr(), None), "bmTraceRays")
    for _ in range(max(args.warmup, 3)):
        step(rays)
    cnt = torch.zeros(4, dtype=torch.int64, device=dev)
    _lib.check(L.bmTraceRays(tree.h, rays.data_ptr(), n, 0, hits.data_ptr(), cnt.data_ptr()), "bmTraceRays")
    torch.cuda.synchronize(dev)
    V, K, Cc, LONG = [int(x) for x in cnt.cpu()]
    hit_rate = float((hits[:, 3].view(torch.int32) != -1).float().mean())
    sampler = ClockSampler(local); launches0 = L.bmLaunchCount()
    sync_all(); sampler.start()
    ms = []
    for _ in range(args.steps):
        a, b = ev(), ev(); sync_all(); a.record(); step(rays); b.record(); sync_all(); ms.append(a.elapsed_time(b))
    clocks = sampler.stop(); launches = L.bmLaunchCount() - launches0
    tsum = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tsum, op=dist.ReduceOp.MAX)
    ms_per_step = float(tsum.item()) / args.steps
    value = n * world / (ms_per_step * 1e-3) / 1e6
    e2e_dev = torch.empty_like(rays); ems = []
    for i in range(2 + args.steps):
        a, b = ev(), ev(); sync_all(); a.record()
        e2e_dev.copy_(rays_pinned, non_blocking=True); step(e2e_dev); hits_host.copy_(hits, non_blocking=True)
        b.record(); sync_all()
        if i >= 2: ems.append(a.elapsed_time(b))
    t2 = torch.tensor([sum(ems)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = n * world / (float(t2.item()) / args.steps * 1e-3) / 1e6
    peak, peak_src = measured_peaks()
    by = n * (24 + 16) + NODE_BYTES * V + CLUSTER_BYTES * Cc + TRI_BYTES * K
    kms = statistics.mean(ms)
    bt = torch.tensor([min(build_ms) if build_ms else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.broadcast(bt, 0)
    bms = float(bt.item())
    build_bytes = 48 * Tn + 12 * Pn + 128 * Nn + 48 * Pn + 48 * Ln + 8 * Pn
    line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c5: synthetic %.0f M-triangle soup (splitmix64 seed 12345), incoherent rays (seed 54321), %.0f M rays per step per GPU, reference semantics" % (args.c5_mtris, args.c5_mrays),
                       "rays_per_step_per_gpu": n, "l2": "tree (%.1f GB) and ray batch (%.1f GB) far exceed L2" % ((Nn * 128 + Pn * 56 + Ln * 48) / 1e9, n * 40 / 1e9),
                       "sharding": "ray batches per rank, tree built on rank 0 and replicated by NCCL broadcast" if world > 1 else "single GPU", "hit_rate": hit_rate},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n * 16, "ms_per_step": float(t2.item()) / args.steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "trace_kernel<rays> (bmTraceRays, reference semantics)", "achieved": by / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": by / (kms * 1e-3) / 1e9 / peak, "traffic": None, "ms_per_launch": kms, "bytes_per_ray": by / n, "node_visits_per_ray": V / n,
                         "tri_tests_per_ray": K / n, "cluster_boxes_per_ray": Cc / n, "longest_ray_steps": LONG,
                         "note": "24 B ray + 16 B hit + 96 B x node/leaf visits + 32 B x cluster boxes + 48 B x triangle tests; " + peak_src},
            "build": {"metric": "BVH build Mtris/s", "value": Tn / (bms * 1e-3) / 1e6 if bms else None, "unit": "Mtris/s", "ms": bms, "triangles": Tn, "pairs": Pn, "leaves": Ln, "nodes": Nn,
                      "roofline": {"bound": "hbm", "achieved": build_bytes / (bms * 1e-3) / 1e9 if bms else None, "peak": peak, "unit": "GB/s",
                                   "frac": build_bytes / (bms * 1e-3) / 1e9 / peak if bms else None, "traffic": None, "algorithmic_bytes": build_bytes},
                      "replicate_ms": rep_ms}}
    sys.stdout.flush(); os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1); dist.barrier(); dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------------
def _oracle_scene(wl):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    from oracle.bindings import OracleLib, OracleScene
    O = OracleLib()
    s = OracleScene(O); s.add_mesh(wl.positions, wl.indices, wl.normals)
    return O, s


def cpu_baseline(wl, initial_rays, bounce_rays, bounce_src, budget_rays=2_100_000, gpu_pixels_fn=None):
    """The CPU restatement of the reference path (oracle/oracle.cpp, kind "port") on the host cores, on a bounded sample of
    the same workload: the whole BVH build, then a strided sample of the primary rays plus the bounce rays they spawned."""
    cores = os.cpu_count() or 1
    O, s = _oracle_scene(wl)
    tb = s.build()
    n = initial_rays.shape[0]
    stride = max(1, n // budget_rays)
    sel = np.arange(0, n, stride)
    r = s.trace_camera(initial_rays[sel], wl.eye, np.eye(3, dtype=np.float32), 0, cores)
    rays, secs = sel.size, r["seconds"]
    if bounce_rays is not None and bounce_rays.shape[0]:
        keep = (bounce_src % stride) == 0
        B = np.ascontiguousarray(bounce_rays[keep])
        if B.shape[0]:
            r2 = s.trace_rays(B[:, :3], B[:, 3:], BOUNCE_MODE, cores)
            rays += B.shape[0]; secs += r2["seconds"]
    s.close()
    out = {"value": rays / secs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port", "cpu": cpu_model(),
           "sample": "every %d-th primary ray (%d) + the bounce rays they spawned, %d rays total, %d threads over rays; tree built on 1 thread" % (stride, sel.size, rays, cores),
           "build": {"value": wl.num_tris / tb / 1e6, "unit": "Mtris/s", "cores": 1, "seconds": tb}}
    ref = reference_build_baseline(wl, gpu_pixels_fn)
    if ref:
        out["reference_build"] = ref
    return out


def reference_build_baseline(wl, gpu_pixels_fn=None, W=960, H=540):
    """The UNMODIFIED reference (oracle/_ref/libbeam_ref.so: its own sources compiled in their CUDA-0 host mode, oracle/build_ref.sh)
    through its public interface on one core -- it is not thread-safe and can only trace a camera's primary rays into pixels, so the
    sample is the workload's scene at W x H primary rays.  When the GPU path's pixels for the same frame are supplied they are compared."""
    try:
        from oracle.bindings import RefLib, RefScene
        if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libbeam_ref.so")):
            return None
        s = RefScene(RefLib())
        s.add_mesh(wl.positions, wl.indices, wl.normals)
        tb = s.build()
        err, px, sec = s.trace(W, H, wl.eye, np.eye(3, dtype=np.float32), wl.frustum)
        s.close()
        if err:
            return {"kind": "reference", "error": int(err)}
        out = {"kind": "reference", "value": W * H / sec / 1e6, "unit": "Mrays/s", "cores": 1,
               "sample": "the workload's scene, %dx%d primary rays through IScene / ICamera::traceScene of the unmodified reference (CUDA 0 host mode)" % (W, H),
               "build": {"value": wl.num_tris / tb / 1e6, "unit": "Mtris/s", "cores": 1, "seconds": tb}}
        if gpu_pixels_fn is not None:
            gp = gpu_pixels_fn(W, H)
            out["gpu_pixels_equal_reference"] = bool(np.array_equal(np.asarray(gp, np.uint32).reshape(-1), np.asarray(px, np.uint32).reshape(-1)))
        return out
    except Exception as ex:                                   # the baseline is reported, never required
        return {"kind": "reference", "unavailable": repr(ex)[:200]}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (its serial host-emulation algorithm as restated in
    oracle/oracle.cpp, validated bit-for-bit against the unmodified reference build oracle/_ref) on all host threads.
    Each step = a bounded sample of the workload's frame (strided primaries + a diffuse bounce for every hit)."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from raytracercuda_b200 import workloads
    wl = workloads.get(args.workload)
    cores = os.cpu_count() or 1
    O, s = _oracle_scene(wl)
    tb = s.build()
    n = wl.width * wl.height
    stride = max(1, n // args.reference_rays)
    rays = O.initial_rays(wl.width, wl.height, wl.frustum)[::stride]
    rng = np.random.default_rng(args.seed)
    eye = np.asarray(wl.eye, np.float32)

    def step():
        r = s.trace_camera(rays, wl.eye, np.eye(3, dtype=np.float32), 0, cores)
        total, secs = rays.shape[0], r["seconds"]
        if wl.bounces:
            hit = r["tri"] != 0xFFFFFFFF
            if hit.any():
                # host-side bounce rays with the same construction (hit + 1e-4 n, cosine-weighted about the geometric normal)
                d = rays[hit]; t = r["t"][hit]
                tri = wl.indices[r["tri"][hit]]
                p0, p1, p2 = wl.positions[tri[:, 0]], wl.positions[tri[:, 1]], wl.positions[tri[:, 2]]
                nrm = np.cross(p1 - p0, p2 - p0); nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)
                nrm = np.where((nrm * d).sum(1, keepdims=True) > 0, -nrm, nrm)
                hp = eye + t[:, None] * d
                u1, u2 = rng.random(t.size), rng.random(t.size)
                rr, phi = np.sqrt(u1), 2 * np.pi * u2
                a = np.where(np.abs(nrm[:, :1]) > 0.9, np.array([[0, 1, 0]]), np.array([[1, 0, 0]]))
                t1 = np.cross(nrm, a); t1 /= np.linalg.norm(t1, axis=1, keepdims=True); t2 = np.cross(nrm, t1)
                nd = t1 * (rr * np.cos(phi))[:, None] + t2 * (rr * np.sin(phi))[:, None] + nrm * np.sqrt(np.maximum(0, 1 - u1))[:, None]
                r2 = s.trace_rays((hp + 1e-4 * nrm).astype(np.float32), nd.astype(np.float32), BOUNCE_MODE, cores)
                total += t.size; secs += r2["seconds"]
        return total, secs

    for _ in range(args.warmup):
        step()
    tot, sec = 0, 0.0
    for _ in range(args.steps):
        a, b = step(); tot += a; sec += b
    value = tot / sec / 1e6
    sample = "every %d-th primary ray (%d) + one diffuse bounce per hit, per step; %d threads over rays" % (stride, rays.shape[0], cores)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "rays_per_step": tot // args.steps, "trace_mode": "reference semantics"},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "cpu": cpu_model(), "sample": sample,
                             "build": {"value": wl.num_tris / tb / 1e6, "unit": "Mtris/s", "cores": 1, "seconds": tb}},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    s.close()
    # Why the port and not oracle/_ref: the unmodified reference is single-threaded (global emulation state) and can only trace a camera's
    # primary rays into pixels, so it cannot run this metric's bounce rays nor use the host's cores; the port is validated bit for bit
    # against it (tests/test_oracle_golden.py) and is the FASTER baseline.  The reference's own single-core number is reported beside it.
    line["cpu_baseline"]["why_port"] = "the unmodified reference is single-threaded and traces camera rays only; the port (bit-exact against it) runs the whole metric on all cores and is the faster baseline"
    ref = reference_build_baseline(wl)
    if ref:
        line["cpu_baseline"]["reference_build"] = ref
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--bands", type=int, default=1, help="row bands of a frame issued concurrently on as many streams (tail filling)")
    ap.add_argument("--no-e2e-graph", action="store_true", help="e2e leg: direct launches instead of one CUDA graph per frame")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c5-mtris", type=float, default=100.0, help="workload c5: millions of soup triangles")
    ap.add_argument("--c5-mrays", type=float, default=16.0, help="workload c5: millions of incoherent rays per step per GPU")
    ap.add_argument("--reference-rays", type=int, default=200_000, help="primary rays per step of the --impl reference sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c5":
        run_c5(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
