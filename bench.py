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


def orbit_pose(eye0, degrees):
    """camera on a circle about the y axis looking at the origin: (eye, orient) with orient in glm's column-major order (m[3*col+row])"""
    a = np.radians(float(degrees)); c, s_ = float(np.cos(a)), float(np.sin(a))
    R = np.array([[c, 0.0, s_], [0.0, 1.0, 0.0], [-s_, 0.0, c]], np.float32)
    return (R @ np.asarray(eye0, np.float32)).astype(np.float32), np.ascontiguousarray(R.T).reshape(-1)


def profile_entry(name):
    """numbers of a committed ncu capture (profiles/traffic.json): DRAM / L2 bytes and warp instructions per launch of a named kernel"""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name)
    except Exception:
        return None


def l2_copy_peak(torch, dev):
    """L2 bandwidth this GPU delivers to a plain copy whose 2 x 24 MiB working set stays in the cache (read + write bytes, best of 20)"""
    a = torch.empty(24 << 20, dtype=torch.uint8, device=dev); b = torch.empty_like(a)
    best = 1e30
    for _ in range(3):
        b.copy_(a)
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize(dev)
        best = min(best, e0.elapsed_time(e1))
    return 2 * a.numel() / (best * 1e-3) / 1e9


def run_b200(args):
    import ctypes as C
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
    SPP = wl.samples
    # weak scaling (default; the driver's 1 -> 8 run): every rank renders a full-size image = interleaved rows of an N-times taller one.
    # strong scaling (--scaling strong; default for the multi-sample c3 / c4 frames): ONE frame of the workload's size; rank r renders the
    # bands of 4 rows r, r + N, ... of it (interleaved image tiles one 8x4 tile row high: every rank sees the same mix of silhouette and
    # background, and its 8x4 tiles are tiles of neighbouring pixels of the real image); the e2e leg, which goes through
    # ICamera::setInitialRays, renders rows r, r + N, ... (the only interleave a frustum can express)
    strong = (args.scaling == "strong") or (args.scaling == "auto" and SPP > 1)
    W = wl.width
    if strong and wl.height % world:
        raise SystemExit("--scaling strong needs the image height to be a multiple of the GPU count")
    H = wl.height // world if strong else wl.height
    n = W * H
    eye0 = np.asarray(wl.eye, np.float32)
    frustum = workloads.interleaved_rows_frustum(wl.frustum, H, rank, world)
    moving = SPP == 1 and args.orbit_deg != 0.0          # the headline frame orbits the scene; multi-sample frames accumulate a static view
    _f3 = lambda a: (C.c_float * 3)(*[float(x) for x in a])
    _f9 = lambda a: (C.c_float * 9)(*[float(x) for x in np.asarray(a, np.float32).reshape(-1)])

    def pose(i, orbit=True):
        e, o = orbit_pose(eye0, args.orbit_deg * i if (orbit and moving) else 0.0)
        return _f3(e), _f9(o), e, o

    # --- scene + camera resident on the device -------------------------------------------------------------------------
    cam = host.ICamera.create()
    assert cam.setInitialRays(W, H, *frustum) == 0
    rays_host = cam.initialRays()
    H_e2e, n_e2e, frustum_h = H, n, H
    if strong and world > 1:
        cam_full = host.ICamera.create()
        assert cam_full.setInitialRays(W, wl.height, *wl.frustum) == 0
        rows = workloads.band_rows(wl.height, rank, world)
        rays_host = np.ascontiguousarray(cam_full.initialRays().reshape(wl.height, W, 3)[rows].reshape(-1, 3))
        H = int(rows.size); n = W * H; frustum_h = wl.height
        del cam_full
    rays_dev = torch.from_numpy(rays_host).to(dev)
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
    stat_keys = ("triangles", "pairs", "leaves", "nodes", "depth", "max_leaf_faces", "overflow_leaves", "meshes")
    stats_t = torch.zeros(8, dtype=torch.int64, device=dev)
    if rank == 0:
        s0 = built.stats()
        stats_t.copy_(torch.tensor([s0[k] for k in stat_keys]))
    if world > 1:
        dist.broadcast(stats_t, 0)
    st = dict(zip(stat_keys, [int(x) for x in stats_t.cpu()]))
    pixels = torch.empty(n, dtype=torch.int32, device=dev)
    # multi-sample frames trace GROUP samples of every pixel per launch, then the bounce rays of those samples in one bounce launch (as ICamera::traceSceneProgressive does)
    GROUP = max(1, min(SPP, 8, (32 << 20) // n))
    hits1 = torch.empty((n * GROUP, 4), dtype=torch.float32, device=dev)
    hits2 = torch.empty((n * GROUP, 4), dtype=torch.float32, device=dev)
    nbounce = torch.zeros(1, dtype=torch.int64, device=dev)              # bounce rays traced in the step (counted by the bounce launches)
    # instrumentation outside the timed region (traversal counters, CPU baseline): the bounce rays of one sample as a list
    rays2 = torch.empty((n, 6), dtype=torch.float32, device=dev)
    src2 = torch.empty(n, dtype=torch.int32, device=dev); cnt2 = torch.zeros(1, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    cnt_host = torch.zeros(1, dtype=torch.int64).pin_memory()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    jit = (abs(wl.frustum[1] - wl.frustum[0]) / W / wl.frustum[4], abs(frustum[3] - frustum[2]) / frustum_h / wl.frustum[4])
    acc = torch.zeros(n, dtype=torch.float32, device=dev) if SPP > 1 else None
    cov = torch.zeros(n, dtype=torch.int32, device=dev) if SPP > 1 else None

    # --- the framebuffer of all ranks on every rank -----------------------------------------------------------------------
    # Preferred: symmetric memory -- every rank writes its rows straight into each peer's copy of the gathered frame with copy-engine
    # transfers over NVLink (no SM is taken from the persistent trace grids); fallback: NCCL all-gather on a side stream.
    gathered, comm, symm = None, None, None
    n_max = n
    if world > 1:
        nm = torch.tensor([n], dtype=torch.int64, device=dev); dist.all_reduce(nm, op=dist.ReduceOp.MAX); n_max = int(nm.item())
        comm = torch.cuda.Stream(dev)
        if not os.environ.get("BENCH_NO_SYMM"):
            try:
                import torch.distributed._symmetric_memory as symm_mem
                gathered = symm_mem.empty((world, n_max), dtype=torch.int32, device=dev)
                hdl = symm_mem.rendezvous(gathered, dist.group.WORLD)
                peers = [hdl.get_buffer(r_, (world, n_max), torch.int32) for r_ in range(world)]
                symm = (hdl, peers)
                # prove the path once against NCCL
                probe = torch.full((n,), rank + 1, dtype=torch.int32, device=dev)
                for r_ in range(world):
                    peers[r_][rank][:n].copy_(probe, non_blocking=True)
                torch.cuda.synchronize(dev); dist.barrier(); torch.cuda.synchronize(dev)
                want = torch.arange(1, world + 1, dtype=torch.int32, device=dev)
                if not bool((gathered[:, 0] == want).all()):
                    raise RuntimeError("symmetric-memory gather check failed")
            except Exception as ex:
                sys.stderr.write("rank %d: symmetric-memory gather unavailable (%r): NCCL all-gather\n" % (rank, ex))
                symm = None
        ok_t = torch.tensor([1 if symm is not None else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        if int(ok_t.item()) == 0:
            symm = None
            gathered = torch.empty((world, n_max), dtype=torch.int32, device=dev)
    pixels_padded = None
    if world > 1 and symm is None and n != n_max:
        pixels_padded = torch.zeros(n_max, dtype=torch.int32, device=dev)     # NCCL all-gather wants equal shares

    def gather_frame(src_pixels, after_event):
        """issue the gather of this rank's rows on the side stream once `after_event` has passed; returns a callable that makes the compute stream wait for it"""
        with torch.cuda.stream(comm):
            comm.wait_event(after_event)
            if symm is not None:
                for k in range(world):
                    r_ = (rank + k) % world
                    symm[1][r_][rank][:n].copy_(src_pixels, non_blocking=True)
                done = torch.cuda.Event(); done.record(comm)
                return lambda: stream.wait_event(done)
            if pixels_padded is not None:
                pixels_padded[:n].copy_(src_pixels); src_pixels = pixels_padded
            work = dist.all_gather_into_tensor(gathered.view(-1), src_pixels, async_op=True)
            return work.wait

    def frame_spp(rays_in, p):
        """multi-sample frame (configs 3 and 4): per sample jittered camera rays -> trace -> bounce rays -> trace -> accumulate; resolve"""
        acc.zero_(); cov.zero_(); nbounce.zero_()
        sp = C.c_void_p(stream.cuda_stream)
        for s0 in range(0, SPP, GROUP):
            g_ = min(GROUP, SPP - s0)
            # g_ jittered samples of every pixel in one camera launch (the jitter is applied where the kernel forms the ray: no 24-byte ray
            # record is written or read), then one bounce launch for the hits of those samples: bounce rays formed in registers, traced as
            # closest hits, accumulated where they end.  Both walk 8x4 tiles under their temporal tile schedules.
            _lib.check(L.bmTraceCameraSamples(tree.h, rays_in.data_ptr(), n, W, p[0], p[1], jit[0], jit[1], s0, g_, args.seed, 0, hits1.data_ptr(), None), "bmTraceCameraSamples")
            _lib.check(L.bmTraceBounceCamera(tree.h, rays_in.data_ptr(), n, W, p[0], p[1], jit[0], jit[1], s0, g_, args.seed, hits1.data_ptr(), s0, 1, args.seed, None, 0.25, 0.75,
                                             acc.data_ptr(), cov.data_ptr(), nbounce.data_ptr(), None), "bmTraceBounceCamera")
        _lib.check(L.bmProgressiveResolve(acc.data_ptr(), cov.data_ptr(), n, SPP, pixels.data_ptr(), sp), "resolve")
        cnt_host.copy_(nbounce, non_blocking=True)
        if world > 1:
            e_ = torch.cuda.Event(); e_.record(stream)
            gather_frame(pixels, e_)()
        return n * SPP

    def frame(rays_in, p, marks=None):
        """one step on the device; returns the primary rays traced by this rank (bounce rays are added by bounce_count())"""
        if SPP > 1:
            if marks is not None:
                for m_ in marks: m_.record()
            return frame_spp(rays_in, p)
        if marks is not None: marks[0].record()
        _lib.check(L.bmTraceCamera(tree.h, rays_in.data_ptr(), n, W, p[0], p[1], 0, hits1.data_ptr(), pixels.data_ptr(), table.data_ptr(), None), "bmTraceCamera")
        if marks is not None: marks[1].record()
        wait = None
        if world > 1:
            # the framebuffer is final once the primary launch is done: gather it on a side stream while the bounce rays are traced
            e_ = torch.cuda.Event(); e_.record(stream)
            wait = gather_frame(pixels, e_)
        if wl.bounces:
            nbounce.zero_()
            if marks is not None: marks[2].record()
            # bounce launch: every pixel with a primary hit forms its diffuse bounce ray in registers and traces it as a closest hit (hit
            # records by pixel in hits2); the ray count stays on the device
            _lib.check(L.bmTraceBounceCamera(tree.h, rays_in.data_ptr(), n, W, p[0], p[1], 0.0, 0.0, 0xFFFFFFFF, 1, 0, hits1.data_ptr(), 0, 1, args.seed, hits2.data_ptr(), 0.0, 0.0,
                                             None, None, nbounce.data_ptr(), None), "bmTraceBounceCamera")
            cnt_host.copy_(nbounce, non_blocking=True)
        if marks is not None: marks[3].record()
        if wait is not None:
            wait()                         # the step ends when the gathered frame is complete
        return n

    def bounce_count():
        stream.synchronize()
        return int(cnt_host[0]) if wl.bounces else 0

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def timed_steps(K, first_frame, orbit):
        """K steps, L2 flushed before each, every step bracketed by a synchronise; returns (ms per step max over ranks, Mrays/s of all ranks, per-launch marks)"""
        step_ms, prim_ms, bounce_ms = [], [], []
        total_rays = 0
        for k in range(K):
            flush.zero_()
            marks = [ev() for _ in range(4)]; e0, e1 = ev(), ev()
            p_k = pose(first_frame + k, orbit)               # the frame's camera pose is an input: prepared before the bracket opens
            sync_all()
            e0.record()
            total_rays += frame(rays_dev, p_k, marks)
            e1.record()
            sync_all()
            total_rays += bounce_count()
            step_ms.append(e0.elapsed_time(e1)); prim_ms.append(marks[0].elapsed_time(marks[1]))
            bounce_ms.append(marks[2].elapsed_time(marks[3]) if wl.bounces else 0.0)
        t_sum = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev); r_sum = torch.tensor([float(total_rays)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_sum, op=dist.ReduceOp.MAX); dist.all_reduce(r_sum, op=dist.ReduceOp.SUM)
        ms = float(t_sum.item()) / K
        return ms, float(r_sum.item()) / K / (ms * 1e-3) / 1e6, statistics.mean(prim_ms), statistics.mean(bounce_ms)

    # --- warm-up -------------------------------------------------------------------------------------------------------
    WU = max(args.warmup, 3)
    for k in range(WU):
        rays_step = frame(rays_dev, pose(k - WU)) + bounce_count()

    # --- timed: K steps, device resident, L2 flushed between steps --------------------------------------------------------
    sampler = ClockSampler(local)
    launches0 = L.bmLaunchCount()
    sync_all()
    sampler.start()
    ms_per_step, value, prim_moving_ms, bounce_moving_ms = timed_steps(args.steps, 0, True)
    clocks = sampler.stop()
    launches = L.bmLaunchCount() - launches0
    static = None
    if moving:
        for k in range(3):
            frame(rays_dev, pose(0, False)); bounce_count()
        s_ms, s_val, s_prim, s_bounce = timed_steps(args.steps, 0, False)
        static = {"value": s_val, "unit": "Mrays/s", "ms_per_step": s_ms, "primary_launch_ms": s_prim, "bounce_launch_ms": s_bounce,
                  "note": "same step with the camera at rest: every frame repeats the previous one, the best case for the temporal tile schedule"}
    # per-ray traversal counters of the static frame (instrumented launches, outside the timed region)
    p0 = pose(0, False)
    for k in range(2):
        rays_static = frame(rays_dev, p0) + bounce_count()
    nb = bounce_count() if SPP == 1 else int(cnt_host[0])
    cnt = torch.zeros(4, dtype=torch.int64, device=dev)
    _lib.check(L.bmTraceCamera(tree.h, rays_dev.data_ptr(), n, W, p0[0], p0[1], 0, hits1.data_ptr(), None, None, cnt.data_ptr()), "bmTraceCamera")
    torch.cuda.synchronize(dev)
    V1, K1, C1, LONG1 = [int(x) for x in cnt.cpu()]
    V2 = K2 = C2 = 0
    nb_last = 0
    def bounce_launch(counters=None):
        """the bounce launch of the pixel-centre frame in hits1 (instrumentation and per-launch timing)"""
        _lib.check(L.bmTraceBounceCamera(tree.h, rays_dev.data_ptr(), n, W, p0[0], p0[1], 0.0, 0.0, 0xFFFFFFFF, 1, 0, hits1.data_ptr(), 0, 1, args.seed, hits2.data_ptr(), 0.0, 0.0,
                                         None, None, nbounce.data_ptr(), counters.data_ptr() if counters is not None else None), "bmTraceBounceCamera")
    if wl.bounces:
        # the same bounce rays as a list (bmBounceRaysCamera) for the CPU baseline, and the bounce launch with traversal counters
        _lib.check(L.bmBounceRaysCamera(tree.h, rays_dev.data_ptr(), p0[0], p0[1], hits1.data_ptr(), n, 0, 1, args.seed, rays2.data_ptr(), src2.data_ptr(), cnt2.data_ptr()), "bmBounceRaysCamera")
        cnt.zero_(); nbounce.zero_()
        bounce_launch(cnt)
        torch.cuda.synchronize(dev)
        nb_last = int(nbounce.item())
        assert nb_last == int(cnt2.item()), "bounce launch and bounce list disagree on the number of rays"
        V2, K2, C2, LONG2 = [int(x) for x in cnt.cpu()]

    # --- timed: e2e through the reference-facing host classes (include/beam_host_c.h) ---------------------------------------------
    # What a Beam program does per frame: ICamera::traceSceneProgressive (c1: traceScene) with this frame's eye / orientation -- the camera's
    # direction table was uploaded ONCE by setInitialRays, as in the reference (Camera.cpp:43-72, 85-97) -- then IRenderTarget::readPixels into
    # pinned host memory.  Two render targets: the read-back of frame i-1 (copy stream) overlaps the kernels of frame i.  The scene was
    # uploaded from HOST arrays through IMesh::setVertexData / setIndices and built by IScene::updateGPUScene.  The K frames are ONE bracket.
    hs = torch.cuda.ExternalStream(int(L.beamHostStream()), device=dev)
    scene = host.IScene.create(); mesh = host.IMesh.create()
    assert mesh.setVertexData(wl.positions, wl.positions.shape[0], 3, host.VERTEX_DATA_POSITION) == 0
    assert mesh.setVertexData(wl.normals, wl.normals.shape[0], 3, host.VERTEX_DATA_NORMAL) == 0
    assert mesh.setIndices(wl.indices.reshape(-1), wl.indices.size) == 0
    scene.addMesh(mesh); scene.updateGPUScene()
    rts = [host.IRenderTarget.createHeadless(W, H_e2e) for _ in range(2)]
    pix_pinned = [torch.empty(n_e2e, dtype=torch.int32).pin_memory() for _ in range(2)]
    rays_pinned = [torch.zeros(1, dtype=torch.int64).pin_memory() for _ in range(2)]
    sticky0 = L.bmStickyError()

    def e2e_frame(i, p):
        rt = rts[i % 2]
        assert rt.lock() == 0
        with torch.cuda.stream(hs):
            flush.zero_()                                  # the L2 flush sits inside the host classes' stream, before the frame's kernels
        if wl.bounces or SPP > 1:
            err = L.beamCameraTraceSceneProgressive(cam.h, p[2].ctypes.data, p[3].ctypes.data, scene.h, SPP, args.seed, C.cast(rays_pinned[i % 2].data_ptr(), C.POINTER(C.c_uint64)))
        else:
            err = L.beamCameraTraceScene(cam.h, p[2].ctypes.data, p[3].ctypes.data, scene.h)
            rays_pinned[i % 2][0] = n_e2e
        assert err == 0, err
        assert rt.unlock() == 0

    def e2e_read(i):
        err = L.beamRenderTargetReadPixels(rts[i % 2].h, pix_pinned[i % 2].data_ptr())
        assert err == 0, err
        return int(rays_pinned[i % 2][0])

    def e2e_run(K, first):
        total = 0
        poses = [pose(first + i) for i in range(K)]
        for i in range(K):
            e2e_frame(i, poses[i])
            if i:
                total += e2e_read(i - 1)
        total += e2e_read(K - 1)
        return total

    e2e_run(3, -3)
    fl = []
    for k in range(5):                                       # what the L2 flush inside every e2e frame costs on its own
        e0, e1 = ev(), ev(); sync_all(); e0.record(hs)
        with torch.cuda.stream(hs):
            flush.zero_()
        e1.record(hs); sync_all(); fl.append(e0.elapsed_time(e1))
    e2e_flush_ms = statistics.mean(fl[1:])
    lat = []
    for k in range(3):                                       # latency of one frame on its own
        e0, e1 = ev(), ev(); sync_all(); e0.record(hs); e2e_run(1, k); e1.record(hs); sync_all(); lat.append(e0.elapsed_time(e1))
    e0, e1 = ev(), ev()
    sync_all()
    e0.record(hs); e2e_rays = e2e_run(args.steps, 0); e1.record(hs)
    sync_all()
    e2e_total_ms = e0.elapsed_time(e1)
    e2e_latency_ms = statistics.mean(lat)
    assert L.bmStickyError() == sticky0
    t2 = torch.tensor([e2e_total_ms], dtype=torch.float64, device=dev); r2 = torch.tensor([float(e2e_rays)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX); dist.all_reduce(r2, op=dist.ReduceOp.SUM)
    e2e_value = float(r2.item()) / (float(t2.item()) * 1e-3) / 1e6
    # the e2e frame of a static pose must show the pixels of the device-timed path (same kernels behind both)
    e2e_frame(0, pose(0, False)); e2e_read(0)
    e2e_pixels = pix_pinned[0].clone()
    h2d = 48
    d2h = n_e2e * 4 + 8
    del scene, mesh

    # --- timed: BVH build (soup resident -> tree traceable) ---------------------------------------------------------------
    build_ms, stages = [], ""
    if rank == 0:
        for i in range(2 + args.steps):
            flush.zero_()
            e0, e1 = ev(), ev()
            torch.cuda.synchronize(dev)
            e0.record(); build_once(); e1.record()
            torch.cuda.synchronize(dev)
            if i >= 2:
                build_ms.append(e0.elapsed_time(e1))
        built.enable_stage_timing(True); build_once(); torch.cuda.synchronize(dev)
        stages = L.bmTreeLastBuildStages(built.h).decode(); built.enable_stage_timing(False)
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
        rt_ = torch.tensor([min(reps)], dtype=torch.float64, device=dev)
        dist.all_reduce(rt_, op=dist.ReduceOp.MAX)
        rep_ms = float(rt_.item())
    bt = torch.tensor([statistics.mean(build_ms) if build_ms else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.broadcast(bt, 0)
    bms = float(bt.item())
    peak, peak_src = measured_peaks()
    node_b = 128 if st["nodes"] * 128 <= 48 * 1048576 else 64
    build_bytes = 48 * st["triangles"] + 12 * st["pairs"] + node_b * st["nodes"] + 48 * st["pairs"] + 48 * st["leaves"] + 8 * st["pairs"]
    build = {"metric": "BVH build Mtris/s", "value": st["triangles"] / (bms * 1e-3) / 1e6, "unit": "Mtris/s", "ms": bms,
             "triangles": st["triangles"], "pairs": st["pairs"], "leaves": st["leaves"], "nodes": st["nodes"], "overflow_leaves": st["overflow_leaves"], "max_leaf_faces": st["max_leaf_faces"],
             "stages_us": stages,
             "roofline": {"bound": "latency" if st["triangles"] < 5_000_000 else "hbm", "achieved": build_bytes / (bms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                          "frac": build_bytes / (bms * 1e-3) / 1e9 / peak, "traffic": None,
                          "algorithmic_bytes": build_bytes, "note": "compulsory bytes of the whole build: 48 B/tri soup read + 12 B/pair (8 B key + 4 B id) written + outputs (%d B/node, 48 B/leaf, 48 B/pair triangle record, 32 B per 4 pairs cluster box); sort passes excluded (SURVEY 8d); at this size the build is a chain of ~30 short launches, bound by launch latency, not by HBM; " % node_b + peak_src},
             "replication": ("built on rank 0, arrays replicated to %d peers by NCCL broadcast in %.3f ms (not included in ms)" % (world - 1, rep_ms)) if world > 1 else "single GPU",
             "replicate_ms": rep_ms}

    # --- roofline of the dominant kernel: the primary-ray trace launch (static frame, the one the committed ncu capture profiled) ----------
    def time_launch(fn, reps=7):
        tt = []
        for _ in range(reps):
            flush.zero_(); e0, e1 = ev(), ev(); torch.cuda.synchronize(dev); e0.record(); fn(); e1.record(); torch.cuda.synchronize(dev); tt.append(e0.elapsed_time(e1))
        return statistics.mean(tt[2:])
    pms = time_launch(lambda: _lib.check(L.bmTraceCamera(tree.h, rays_dev.data_ptr(), n, W, p0[0], p0[1], 0, hits1.data_ptr(), pixels.data_ptr(), table.data_ptr(), None), "bmTraceCamera"))
    prim_bytes = n * (12 + 16 + 4) + (NODE_BYTES if node_b == 128 else 64) * V1 + TRI_BYTES * K1 + CLUSTER_BYTES * C1
    sm_clock_ghz = (clocks.get("sm_mhz") or 1965.0) / 1e3
    issue_peak = 148 * 4 * sm_clock_ghz                       # warp instructions per ns = G inst/s: 4 schedulers per SM, one issue per clock each
    prof = profile_entry(args.workload) if (world == 1 or not strong) else None
    tree_mb = (st["nodes"] * node_b + st["pairs"] * 56 + st["leaves"] * 48) / 1e6
    roofline = {"kernel": "trace_kernel<camera> (bmTraceCamera, reference semantics), static frame", "ms_per_launch": pms, "mrays_per_s": n / (pms * 1e-3) / 1e6,
                "algorithmic_bytes_per_ray": prim_bytes / n, "algorithmic_gbs": prim_bytes / (pms * 1e-3) / 1e9,
                "node_visits_per_ray": V1 / n, "tri_tests_per_ray": K1 / n, "cluster_boxes_per_ray": C1 / n, "longest_ray_steps": LONG1,
                "hbm_peak_gbs": peak, "hbm_peak_source": peak_src}
    if prof:
        # bound = the resource the launch uses the largest share of, both shares from the committed ncu capture of this very launch
        # (warp instructions, DRAM and L2 bytes per launch) over the duration measured live
        l2_peak = l2_copy_peak(torch, dev)
        inst = prof["warp_instructions_per_launch"]
        issue_frac = inst / (pms * 1e-3) / 1e9 / issue_peak
        dram_gbs = prof["dram_bytes_per_launch"] / (pms * 1e-3) / 1e9
        hbm_bound = dram_gbs / peak > issue_frac
        roofline.update({"bound": "hbm" if hbm_bound else "issue",
                         "achieved": dram_gbs if hbm_bound else inst / (pms * 1e-3) / 1e9, "peak": peak if hbm_bound else issue_peak, "unit": "GB/s" if hbm_bound else "Gwarp-inst/s",
                         "frac": dram_gbs / peak if hbm_bound else issue_frac,
                         "traffic": prof["dram_bytes_per_launch"], "traffic_source": prof["source"],
                         "issue": {"achieved_gwarp_inst_per_s": inst / (pms * 1e-3) / 1e9, "peak": issue_peak, "frac": issue_frac},
                         "dram": {"achieved_gbs": dram_gbs, "frac_of_hbm_peak": dram_gbs / peak},
                         "l2": {"lts_bytes_per_launch": prof["lts_bytes_per_launch"], "achieved_gbs": prof["lts_bytes_per_launch"] / (pms * 1e-3) / 1e9, "copy_peak_gbs": l2_peak,
                                "frac": prof["lts_bytes_per_launch"] / (pms * 1e-3) / 1e9 / l2_peak, "peak_source": "live: torch copy of two L2-resident 24 MiB buffers (read + write bytes)",
                                "ncu_pct_of_peak": prof.get("lts_pct_of_peak")},
                         "lanes_per_instruction": prof.get("lanes_per_instruction"), "ipc": prof.get("ipc"),
                         "note": "tree %.0f MB; issue peak = 148 SMs x 4 schedulers x %.3f GHz; warp instructions and L2 / DRAM bytes per launch from the committed ncu capture of this very launch, duration measured live; "
                                 "the algorithmic bytes per ray are mostly served from L1 / L2, which is why algorithmic_gbs may exceed the HBM peak" % (tree_mb, sm_clock_ghz)})
    else:
        # no committed capture of this configuration (e.g. a rank's share of a strong-scaling frame): no fraction is claimed
        roofline.update({"bound": "unprofiled", "achieved": None, "peak": None, "unit": None, "frac": None, "traffic": None,
                         "note": "no ncu capture of this configuration is committed, so no roofline fraction is claimed; algorithmic_gbs = (96 or 64 B x node visits + 32 B x cluster boxes + "
                                 "48 B x triangle tests + ray / hit / pixel records) / live duration, mostly served from L1 / L2 (tree %.0f MB); " % tree_mb + peak_src})
    if wl.bounces and nb_last:
        _lib.check(L.bmTraceCamera(tree.h, rays_dev.data_ptr(), n, W, p0[0], p0[1], 0, hits1.data_ptr(), None, None, None), "bmTraceCamera")
        bms2 = time_launch(lambda: bounce_launch())
        roofline["bounce_launch"] = {"kernel": "trace_kernel<closest, camera, bounce> (bmTraceBounceCamera), static frame", "ms_per_launch": bms2, "mrays_per_s": nb_last / (bms2 * 1e-3) / 1e6, "rays": nb_last, "mode": "closest hit",
                                     "node_visits_per_ray": V2 / nb_last, "tri_tests_per_ray": K2 / nb_last, "cluster_boxes_per_ray": C2 / nb_last}
        pb = profile_entry(args.workload + "_bounce")
        if pb and world == 1:
            roofline["bounce_launch"].update({"bound": "issue", "frac": pb["warp_instructions_per_launch"] / (bms2 * 1e-3) / 1e9 / issue_peak, "traffic": pb["dram_bytes_per_launch"],
                                              "lanes_per_instruction": pb.get("lanes_per_instruction"), "traffic_source": pb["source"]})
    # the metric's name says closest hit: the same primary launch in BM_TRACE_CLOSEST (exhaustive, t >= 0) beside the reference-semantics number
    hits_c = torch.empty_like(hits1)
    cms = time_launch(lambda: _lib.check(L.bmTraceCamera(tree.h, rays_dev.data_ptr(), n, W, p0[0], p0[1], 1, hits_c.data_ptr(), None, None, None), "bmTraceCamera"))
    _lib.check(L.bmTraceCamera(tree.h, rays_dev.data_ptr(), n, W, p0[0], p0[1], 0, hits1.data_ptr(), pixels.data_ptr(), table.data_ptr(), None), "bmTraceCamera")
    torch.cuda.synchronize(dev)
    same = float((hits_c[:, 3].view(torch.int32) == hits1[:, 3].view(torch.int32)).float().mean())
    closest_mode = {"primary_launch_ms": cms, "primary_mrays_per_s": n / (cms * 1e-3) / 1e6, "reference_mode_primary_launch_ms": pms,
                    "same_triangle_as_reference_mode": same,
                    "note": "BM_TRACE_CLOSEST = exhaustive closest hit with t >= 0; the reference's own semantics (first leaf with a hit, any sign of t) is the parity mode the primary rays are timed in; bounce rays are always traced in closest mode"}
    e2e_equal = bool(torch.equal(e2e_pixels.to(dev), pixels)) if (SPP == 1 and not wl.bounces) else None

    line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": WU,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "rays_per_step_per_gpu": rays_step, "primary_rays_per_gpu": n * SPP, "bounce_rays_per_gpu": nb, "samples_per_pixel": SPP,
                       "image_per_gpu": [W, H],
                       "trace_mode": "primary rays: reference semantics (BM_TRACE_REFERENCE); bounce rays: closest hit, t >= 0 (BM_TRACE_CLOSEST)", "l2": "flushed between timed steps (256 MiB write)", "cpu_affinity": numa,
                       "sharding": ("interleaved image rows per rank, replicated tree, framebuffer of all ranks assembled on every rank by %s, overlapped with the bounce trace"
                                    % ("copy-engine writes into the peers' symmetric memory over NVLink" if symm is not None else "NCCL all-gather")) if world > 1 else "single GPU",
                       "camera": {"eye": [float(x) for x in eye0], "frustum": list(wl.frustum),
                                  "motion": ("orbit about the y axis, %.3g degrees per frame, looking at the origin (every frame differs from the one the tile schedule learnt from)" % args.orbit_deg) if moving else "static"}},
            "primary_launch_ms": prim_moving_ms, "bounce_launch_ms": bounce_moving_ms,
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": float(t2.item()) / args.steps,
                    "single_frame_latency_ms": e2e_latency_ms, "l2_flush_ms_inside_every_frame": e2e_flush_ms, "pixels_equal_device_path": e2e_equal,
                    "api": "host classes through include/beam_host_c.h: " + ("beamCameraTraceSceneProgressive" if (wl.bounces or SPP > 1) else "beamCameraTraceScene") + " + beamRenderTargetReadPixels; scene uploaded from host arrays by beamMeshSetVertexData / beamMeshSetIndices, built by beamSceneUpdateGPUScene",
                    "pipeline": "per frame: eye + orientation (48 B, kernel arguments; the camera's direction table was uploaded once by setInitialRays, as the reference does) -> frame on the host classes' stream -> "
                                "read-back of the framebuffer and the ray count into pinned memory on the copy stream; two render targets, so the read-back of frame i-1 overlaps frame i; K frames = one timed bracket, "
                                "which therefore also contains one 256 MiB L2 flush per frame (l2_flush_ms_inside_every_frame; the device-timed step flushes outside its bracket)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "build": build, "closest_mode": closest_mode}
    if static is not None:
        line["static_camera"] = static

    if rank == 0 and world == 1 and not args.no_hbm_leg and args.workload == "c2":
        del flush
        torch.cuda.empty_cache()
        try:
            line["hbm_leg"] = hbm_leg(torch, dev, L, args)
        except Exception as ex:                               # reported, never fatal for the headline
            line["hbm_leg"] = {"error": repr(ex)[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if SPP == 1 and wl.bounces:
            _lib.check(L.bmTraceCamera(tree.h, rays_dev.data_ptr(), n, W, p0[0], p0[1], 0, hits1.data_ptr(), None, None, None), "bmTraceCamera")
            _lib.check(L.bmBounceRaysCamera(tree.h, rays_dev.data_ptr(), p0[0], p0[1], hits1.data_ptr(), n, 0, 1, args.seed, rays2.data_ptr(), src2.data_ptr(), cnt2.data_ptr()), "bmBounceRaysCamera")
            torch.cuda.synchronize(dev)
            nb = int(cnt2.item())
        def gpu_frame_pixels(w_, h_):
            """the product's pixels for the workload's scene at w_ x h_ (same camera), for the comparison with the reference build"""
            c_ = host.ICamera.create(); assert c_.setInitialRays(w_, h_, *wl.frustum) == 0
            r_ = torch.from_numpy(c_.initialRays()).to(dev); p_ = torch.empty(w_ * h_, dtype=torch.int32, device=dev)
            _lib.check(L.bmTraceCamera(tree.h, r_.data_ptr(), w_ * h_, w_, p0[0], p0[1], 0, None, p_.data_ptr(), table.data_ptr(), None), "bmTraceCamera")
            torch.cuda.synchronize(dev)
            return p_.cpu().numpy().view(np.uint32)
        nbb = nb if SPP == 1 else nb_last                    # multi-sample frames: the bounce list of the pixel-centre sample
        line["cpu_baseline"] = cpu_baseline(wl, rays_host, rays2[:nbb].cpu().numpy() if nbb else None, src2[:nbb].cpu().numpy() if nbb else None,
                                            gpu_pixels_fn=gpu_frame_pixels)
        def ours_scene(name_):
            """this library on one of the reference-kernel scenes: build (best of 3) and the primary launch at 1280x720 (mean of 5 after warm-up)"""
            P_, I_, N_ = workloads._bundled(name_) if name_ in ("f16", "suzanne", "bunny") else (wl.positions, wl.indices, wl.normals)
            Pd_ = torch.from_numpy(P_).to(dev); Id_ = torch.from_numpy(I_.view(np.int32)).to(dev).reshape(-1); Nd_ = torch.from_numpy(N_).to(dev)
            tab_ = make_mesh_table([(Pd_, Id_, Nd_)], dev)
            t_ = Tree(dev); bt_ = []
            for _ in range(4):
                t_.reset(); t_.add_mesh(Pd_, Id_); e0, e1 = ev(), ev(); torch.cuda.synchronize(dev); e0.record(); t_.build(); e1.record(); torch.cuda.synchronize(dev); bt_.append(e0.elapsed_time(e1))
            c_ = host.ICamera.create(); assert c_.setInitialRays(1280, 720, *wl.frustum) == 0
            r_ = torch.from_numpy(c_.initialRays()).to(dev); p_ = torch.empty(1280 * 720, dtype=torch.int32, device=dev)
            e_ = _f3([0.0, 0.0, -2.1]); o_ = _f9(np.eye(3, dtype=np.float32)); tt_ = []
            for _ in range(8):
                e0, e1 = ev(), ev(); torch.cuda.synchronize(dev); e0.record()
                _lib.check(L.bmTraceCamera(t_.h, r_.data_ptr(), 1280 * 720, 1280, e_, o_, 0, None, p_.data_ptr(), tab_.data_ptr(), None), "bmTraceCamera")
                e1.record(); torch.cuda.synchronize(dev); tt_.append(e0.elapsed_time(e1))
            t_.close()
            return {"build_ms": min(bt_[1:]), "build_mtris_per_s": I_.shape[0] / (min(bt_[1:]) * 1e-3) / 1e6, "trace_ms": statistics.mean(tt_[3:]),
                    "primary_mrays_per_s": 1280 * 720 / (statistics.mean(tt_[3:]) * 1e-3) / 1e6}
        line["reference_cuda_b200"] = reference_cuda_baseline(args.workload, ours_scene)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        os.dup2(2, 1)
        dist.barrier()
        dist.destroy_process_group()


def hbm_leg(torch, dev, L, args):
    """The memory-bound side of the path inside the default line: a 20 M-triangle soup (tree ~6 GB, 50 x the L2) and 16 M incoherent rays --
    BASELINE config 5 at a size that builds and traces in about two seconds.  Both traversal forms are timed: the direct one (one DRAM round
    trip per node: the HBM-bound kernel, its fraction from the DRAM bytes of the committed ncu capture) and the brick-by-brick one bmTraceRays
    chooses for such trees (same hits, the bytes mostly served by the L2: issue-bound)."""
    from raytracercuda_b200 import _lib, synth
    from raytracercuda_b200.tree import Tree
    T, n = int(args.hbm_mtris * 1e6), int(args.hbm_mrays * 1e6)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    Pd, Id = synth.soup_torch(T, dev)
    t = Tree(dev)
    bms = []
    for i in range(3):
        t.reset(); t.add_mesh(Pd, Id)
        e0, e1 = ev(), ev(); torch.cuda.synchronize(dev); e0.record(); t.build(); e1.record(); torch.cuda.synchronize(dev); bms.append(e0.elapsed_time(e1))
    t.enable_stage_timing(True); t.reset(); t.add_mesh(Pd, Id); t.build(); torch.cuda.synchronize(dev)
    stages = L.bmTreeLastBuildStages(t.h).decode(); t.enable_stage_timing(False)
    st = t.stats()
    R = synth.incoherent_rays_torch(n, dev)
    hits = torch.empty((n, 4), dtype=torch.float32, device=dev); ref_hits = torch.empty_like(hits)
    peak, peak_src = measured_peaks()
    out = {"workload": "c5-class: synthetic %.0f M-triangle soup, %.0f M incoherent rays per launch (origin uniform in the world, direction uniform on the sphere)" % (args.hbm_mtris, args.hbm_mrays),
           "tree_gb": (st["nodes"] * 64 + st["pairs"] * 56 + st["leaves"] * 48) / 1e9, "triangles": st["triangles"], "pairs": st["pairs"], "leaves": st["leaves"],
           "build": {"value": T / (min(bms) * 1e-3) / 1e6, "unit": "Mtris/s", "ms": min(bms), "stages_us": stages}}
    for bricks, key in ((0, "direct"), (1, "bricked")):
        t.set_bricks(bricks, 0)
        leg = {}
        for mode, mname in ((0, "reference"), (1, "closest")):
            ms = []
            for r in range(4):
                e0, e1 = ev(), ev(); torch.cuda.synchronize(dev); e0.record()
                _lib.check(L.bmTraceRays(t.h, R.data_ptr(), n, mode, hits.data_ptr(), None), "bmTraceRays")
                e1.record(); torch.cuda.synchronize(dev); ms.append(e0.elapsed_time(e1))
            m_ = statistics.mean(ms[1:])
            leg[mname] = {"ms_per_launch": m_, "mrays_per_s": n / (m_ * 1e-3) / 1e6}
            if bricks == 0 and mode == 0:
                ref_hits.copy_(hits)
            if bricks == 1 and mode == 0:
                leg["hits_equal_direct"] = bool(torch.equal(hits.view(torch.int32), ref_hits.view(torch.int32)))      # bit patterns: a miss carries tri = 0xFFFFFFFF, a NaN as float
            if mode == 0:
                cnt = torch.zeros(4, dtype=torch.int64, device=dev)
                _lib.check(L.bmTraceRays(t.h, R.data_ptr(), n, mode, hits.data_ptr(), cnt.data_ptr()), "bmTraceRays")
                torch.cuda.synchronize(dev)
                V, K, Cc, LONG = [int(x) for x in cnt.cpu()]
                leg["node_visits_per_ray"] = V / n; leg["tri_tests_per_ray"] = K / n
                leg["algorithmic_bytes_per_ray"] = 40 + 64 * V / n + 48 * K / n
        prof = profile_entry("hbm_leg_" + key)
        ms0 = leg["reference"]["ms_per_launch"]
        if prof:
            scale = n / prof["rays_per_launch"]               # the capture traced rays_per_launch rays of the same batch: bytes and instructions scale with the ray count
            dram = prof["dram_bytes_per_launch"] * scale
            issue_peak = 148 * 4 * 1.965
            issue_frac = prof["warp_instructions_per_launch"] * scale / (ms0 * 1e-3) / 1e9 / issue_peak
            hbm_frac = dram / (ms0 * 1e-3) / 1e9 / peak
            leg["roofline"] = {"bound": "hbm" if hbm_frac > issue_frac else "issue", "frac": max(hbm_frac, issue_frac),
                               "traffic": dram, "traffic_source": prof["source"], "dram_gbs": dram / (ms0 * 1e-3) / 1e9,
                               "frac_of_hbm_peak": hbm_frac, "peak": peak, "frac_of_issue_peak": issue_frac, "issue_peak_gwarp_inst_per_s": issue_peak,
                               "ipc": prof.get("ipc"), "lanes_per_instruction": prof.get("lanes_per_instruction"), "long_scoreboard_stall": prof.get("long_scoreboard_stall"),
                               "note": "reference-mode launch; DRAM bytes and warp instructions of the trace kernel in the committed capture scaled to this launch's ray count, duration of the whole call measured live; bound = the larger share; " + peak_src}
        else:
            a = leg["algorithmic_bytes_per_ray"] * n / (ms0 * 1e-3) / 1e9
            leg["roofline"] = {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": None, "note": "algorithmic bytes (64 B x node visits + 48 B x triangle tests + 40 B ray / hit), no capture committed; " + peak_src}
        out[key] = leg
    out["value"] = out["bricked"]["reference"]["mrays_per_s"]; out["unit"] = "Mrays/s"
    out["value_note"] = "reference semantics through bmTraceRays as shipped (brick by brick for this tree size); closest-hit and direct-traversal numbers beside it"
    t.close()
    return out


def reference_cuda_baseline(workload, ours=None):
    """Second baseline (north_star): the reference's OWN CUDA kernels (BuildTree.cu / RTClear.cu) rebuilt for sm_100a by oracle/build_ref_cuda.sh, run
    per scene in a process of their own (their pool allocation is racy, BuildTree.cu:13-18,43-50, and dies on larger meshes).  Part of the
    cpu_baseline leg: the only place bench.py may execute anything under oracle/."""
    so = os.path.join(ROOT, "oracle", "_ref", "libbeam_refcuda.so")
    if not os.path.exists(so):
        return {"unavailable": "oracle/_ref/libbeam_refcuda.so not built (needs /root/reference at build time)"}
    out = {"what": "Raytracer/BuildTree.cu + RTClear.cu of the reference compiled for sm_100a (--use_fast_math as shipped), primary rays at 1280x720 through its own host classes"}
    for name in ("f16", "suzanne", "bunny", workload):
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_refcuda.py"), name, "1280", "720", "3"], capture_output=True, text=True, timeout=180)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode == 0 and line:
                d = json.loads(line[-1])
                out[name] = {"triangles": d["triangles"], "build_mtris_per_s": d["build_mtris_per_s"], "primary_mrays_per_s": d["primary_mrays_per_s"], "build_ms": d["build_ms"], "trace_ms": d["trace_ms"]}
            else:
                out[name] = "crashed (exit %d: %s)" % (r.returncode, (r.stderr.strip().splitlines() or ["no message"])[-1][:160])
        except subprocess.TimeoutExpired:
            out[name] = "timed out (180 s)"
        except Exception as ex:
            out[name] = "unavailable (%r)" % (ex,)
        if ours is not None:
            try:
                out["this_library_" + name] = ours(name)
            except Exception as ex:
                out["this_library_" + name] = "unavailable (%r)" % (ex,)
    return out


# ---------------------------------------------------------------------------------------------------------------------------
def run_c5(args):
    """BASELINE config 5: counter-based triangle soup (default 100 M triangles) built once on rank 0 and replicated, then ONE fixed batch of
    incoherent rays (origin uniform in the world, direction uniform on the sphere; --c5-total-mrays, default 256 M, generated on the device)
    split by range over the ranks -- strong scaling.  One step = the whole batch, traced in launches of --c5-mrays million rays per rank,
    in the reference's semantics through bmTraceRays as shipped (brick by brick for a tree of this size); the closest-hit mode and the direct
    traversal are timed beside it on one launch."""
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
    T = int(args.c5_mtris * 1e6); chunk = int(args.c5_mrays * 1e6)
    total = int(args.c5_total_mrays * 1e6)
    total -= total % (world * chunk) if total >= world * chunk else 0
    if total < world * chunk:
        total = world * chunk
    mine = total // world                                   # rays [rank * mine, (rank + 1) * mine) of the batch
    nchunks = mine // chunk
    ev = lambda: torch.cuda.Event(enable_timing=True)
    built, build_ms, st, stages = None, [], None, ""
    if rank == 0:
        Pd, Id = synth.soup_torch(T, dev)
        built = Tree(dev)
        for i in range(3):
            built.reset(); built.add_mesh(Pd, Id)
            e0, e1 = ev(), ev(); torch.cuda.synchronize(dev); e0.record(); built.build(); e1.record(); torch.cuda.synchronize(dev)
            build_ms.append(e0.elapsed_time(e1))
        built.enable_stage_timing(True); built.reset(); built.add_mesh(Pd, Id); built.build(); torch.cuda.synchronize(dev)
        stages = L.bmTreeLastBuildStages(built.h).decode(); built.enable_stage_timing(False)
        st = built.stats()
    e0, e1 = ev(), ev(); torch.cuda.synchronize(dev)
    if world > 1: dist.barrier()
    e0.record(); tree = replicate_tree(built, dev, 0); e1.record(); torch.cuda.synchronize(dev)
    rep_t = torch.tensor([e0.elapsed_time(e1) if world > 1 else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(rep_t, op=dist.ReduceOp.MAX)
    rep_ms = float(rep_t.item())
    stt = torch.zeros(4, dtype=torch.int64, device=dev)
    if rank == 0:
        stt.copy_(torch.tensor([st["triangles"], st["pairs"], st["leaves"], st["nodes"]]))
    if world > 1:
        dist.broadcast(stt, 0)
    Tn, Pn, Ln, Nn = [int(x) for x in stt.cpu()]
    # the rank's share of the batch, two chunk buffers: chunk k+1 is generated (counter-based, on the device) while ... no: generation is
    # outside the timed region -- the share is materialised once if it fits (24 B per ray), else regenerated per chunk before the step
    rays_all = [synth.incoherent_rays_torch(chunk, dev, start=rank * mine + k * chunk) for k in range(nchunks)]
    hits = torch.empty((chunk, 4), dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream(dev)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier(); torch.cuda.synchronize(dev)

    hit_count = torch.zeros(1, dtype=torch.int64, device=dev)

    def step(mode=0, count=False):
        for k in range(nchunks):
            _lib.check(L.bmTraceRays(tree.h, rays_all[k].data_ptr(), chunk, mode, hits.data_ptr(), None), "bmTraceRays")
            if count:
                hit_count.add_((hits[:, 3].view(torch.int32) != -1).sum())
    for _ in range(max(min(args.warmup, 1), 1)):
        step(count=True)
    torch.cuda.synchronize(dev)
    hit_rate = float(hit_count.item()) / mine
    sampler = ClockSampler(local); launches0 = L.bmLaunchCount()
    sync_all(); sampler.start()
    ms = []
    for _ in range(args.steps):
        a, b = ev(), ev(); sync_all(); a.record(); step(); b.record(); sync_all(); ms.append(a.elapsed_time(b))
    clocks = sampler.stop(); launches = L.bmLaunchCount() - launches0
    tsum = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tsum, op=dist.ReduceOp.MAX)
    ms_per_step = float(tsum.item()) / args.steps
    value = total / (ms_per_step * 1e-3) / 1e6

    # one launch on its own: per mode, as shipped (bricks chosen by the library) and with the direct traversal; counters of the traversal
    R0 = rays_all[0]
    def one(mode, bricks):
        _lib.check(L.bmTreeSetBricks(tree.h, bricks, 0), "bmTreeSetBricks")
        tt = []
        for r in range(3):
            a, b = ev(), ev(); torch.cuda.synchronize(dev); a.record()
            _lib.check(L.bmTraceRays(tree.h, R0.data_ptr(), chunk, mode, hits.data_ptr(), None), "bmTraceRays")
            b.record(); torch.cuda.synchronize(dev); tt.append(a.elapsed_time(b))
        m_ = statistics.mean(tt[1:])
        return {"ms_per_launch": m_, "mrays_per_s": chunk / (m_ * 1e-3) / 1e6}
    legs = {"as_shipped": {"reference": one(0, -1), "closest": one(1, -1)}, "direct": {"reference": one(0, 0), "closest": one(1, 0)}}
    _lib.check(L.bmTreeSetBricks(tree.h, -1, 0), "bmTreeSetBricks")
    cnt = torch.zeros(4, dtype=torch.int64, device=dev)
    _lib.check(L.bmTraceRays(tree.h, R0.data_ptr(), chunk, 0, hits.data_ptr(), cnt.data_ptr()), "bmTraceRays")
    torch.cuda.synchronize(dev)
    V, K, Cc, LONG = [int(x) for x in cnt.cpu()]

    # e2e: one launch's rays from pinned host memory, its hit records back to pinned host memory
    rays_pinned = R0.cpu().pin_memory(); hits_host = torch.empty((chunk, 4), dtype=torch.float32).pin_memory()
    e2e_dev = torch.empty_like(R0); ems = []
    for i in range(2 + min(args.steps, 5)):
        a, b = ev(), ev(); sync_all(); a.record()
        e2e_dev.copy_(rays_pinned, non_blocking=True)
        _lib.check(L.bmTraceRays(tree.h, e2e_dev.data_ptr(), chunk, 0, hits.data_ptr(), None), "bmTraceRays")
        hits_host.copy_(hits, non_blocking=True)
        b.record(); sync_all()
        if i >= 2: ems.append(a.elapsed_time(b))
    t2 = torch.tensor([statistics.mean(ems)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = chunk * world / (float(t2.item()) * 1e-3) / 1e6
    peak, peak_src = measured_peaks()
    by = chunk * (24 + 16) + 64 * V + CLUSTER_BYTES * Cc + TRI_BYTES * K
    kms = legs["as_shipped"]["reference"]["ms_per_launch"]
    bt = torch.tensor([min(build_ms) if build_ms else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.broadcast(bt, 0)
    bms = float(bt.item())
    build_bytes = 48 * Tn + 12 * Pn + 64 * Nn + 48 * Pn + 48 * Ln + 8 * Pn
    prof = profile_entry("c5")
    roof = {"kernel": "bmTraceRays, reference semantics, one launch of %d rays as shipped (segments + sort + trace_kernel<SEG> + resolve)" % chunk,
            "ms_per_launch": kms, "algorithmic_bytes_per_ray": by / chunk, "algorithmic_gbs": by / (kms * 1e-3) / 1e9, "node_visits_per_ray": V / chunk, "tri_tests_per_ray": K / chunk,
            "longest_segment_steps": LONG}
    if prof:
        dram = prof["dram_bytes_per_launch"] * chunk / prof["rays_per_launch"]
        issue_peak = 148 * 4 * ((clocks.get("sm_mhz") or 1965.0) / 1e3)
        inst_rate = prof["warp_instructions_per_launch"] * chunk / prof["rays_per_launch"] / (kms * 1e-3) / 1e9
        hbm_frac, issue_frac = dram / (kms * 1e-3) / 1e9 / peak, inst_rate / issue_peak
        hb = hbm_frac > issue_frac
        roof.update({"bound": "hbm" if hb else "issue", "achieved": dram / (kms * 1e-3) / 1e9 if hb else inst_rate, "peak": peak if hb else issue_peak, "unit": "GB/s" if hb else "Gwarp-inst/s",
                     "frac": max(hbm_frac, issue_frac), "traffic": dram, "traffic_source": prof["source"],
                     "dram": {"achieved_gbs": dram / (kms * 1e-3) / 1e9, "frac_of_hbm_peak": hbm_frac}, "issue": {"achieved_gwarp_inst_per_s": inst_rate, "peak": issue_peak, "frac": issue_frac},
                     "ipc": prof.get("ipc"), "lanes_per_instruction": prof.get("lanes_per_instruction"),
                     "note": "DRAM bytes and warp instructions of trace_kernel<SEG> in the committed ncu capture (scaled by ray count) / live duration of the whole call; bound = the larger share: the bricked traversal "
                             "is bound by instruction issue (~500 thread instructions per 4-wide node visit), its DRAM traffic is a third of the direct traversal's; " + peak_src})
    else:
        roof.update({"bound": "hbm", "achieved": by / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": by / (kms * 1e-3) / 1e9 / peak, "traffic": None,
                     "note": "algorithmic bytes: 24 B ray + 16 B hit + 64 B x node visits + 48 B x triangle tests; " + peak_src})
    line = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": 1, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "c5: synthetic %.0f M-triangle soup (splitmix64 seed 12345), one batch of %.0f M incoherent rays (seed 54321) split by range over the ranks, %.0f M rays per launch, reference semantics" % (args.c5_mtris, total / 1e6, args.c5_mrays),
                       "rays_per_step": total, "rays_per_step_per_gpu": mine, "launches_per_step_per_gpu": nchunks,
                       "l2": "tree (%.1f GB) and ray batch far exceed L2" % ((Nn * 64 + Pn * 56 + Ln * 48) / 1e9),
                       "sharding": "ray batch split by range, tree built on rank 0 and replicated by NCCL broadcast" if world > 1 else "single GPU", "hit_rate": hit_rate},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": chunk * 24, "d2h_bytes_per_step": chunk * 16, "ms_per_step": float(t2.item()),
                    "note": "one launch per rank: rays from pinned host memory, hit records back to pinned host memory"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "one_launch": legs,
            "build": {"metric": "BVH build Mtris/s", "value": Tn / (bms * 1e-3) / 1e6 if bms else None, "unit": "Mtris/s", "ms": bms, "triangles": Tn, "pairs": Pn, "leaves": Ln, "nodes": Nn,
                      "stages_us": stages,
                      "roofline": {"bound": "hbm", "achieved": build_bytes / (bms * 1e-3) / 1e9 if bms else None, "peak": peak, "unit": "GB/s",
                                   "frac": build_bytes / (bms * 1e-3) / 1e9 / peak if bms else None, "traffic": None, "algorithmic_bytes": build_bytes},
                      "replicate_ms": rep_ms, "replication": "NCCL broadcast of %.1f GB of arrays to %d peers" % ((Nn * 64 + Pn * 56 + Ln * 48 + Tn * 48 + (Pn / 4 + Ln) * 32) / 1e9, world - 1) if world > 1 else "single GPU"}}
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
    # in a process of its own: the reference's fixed-size pools overflow on scenes beyond what it was written for (a segmentation fault,
    # not an error code), and a crash is a result to report, not a reason to lose the bench line
    try:
        if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libbeam_ref.so")):
            return None
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            npy = os.path.join(td, "pixels.npy")
            r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_ref_build.py"), wl.key, str(W), str(H), npy], capture_output=True, text=True, timeout=900)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")]
            if r.returncode != 0 or not line:
                return {"kind": "reference", "crashed": "exit %d%s" % (r.returncode, " (signal %d: its fixed-size pools overflow on this scene)" % -r.returncode if r.returncode < 0 else "")}
            d = json.loads(line[-1])
            if d["error"]:
                return {"kind": "reference", "error": int(d["error"])}
            px = np.load(npy)
        tb, sec = d["build_seconds"], d["trace_seconds"]
        out = {"kind": "reference", "value": W * H / sec / 1e6, "unit": "Mrays/s", "cores": 1,
               "sample": "the workload's scene, %dx%d primary rays through IScene / ICamera::traceScene of the unmodified reference (CUDA 0 host mode), in a process of its own" % (W, H),
               "build": {"value": wl.num_tris / tb / 1e6, "unit": "Mtris/s", "cores": 1, "seconds": tb}}
        if gpu_pixels_fn is not None:
            gp = gpu_pixels_fn(W, H)
            out["gpu_pixels_equal_reference"] = bool(np.array_equal(np.asarray(gp, np.uint32).reshape(-1), np.asarray(px, np.uint32).reshape(-1)))
        return out
    except subprocess.TimeoutExpired:
        return {"kind": "reference", "unavailable": "timed out (900 s)"}
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
    ap.add_argument("--scaling", default="auto", choices=["auto", "weak", "strong"], help="N > 1: weak = every rank renders a full-size image (rows of an N-times taller one); strong = one frame of the workload's size split by rows; auto = weak for c1 / c2, strong for the multi-sample c3 / c4 frames and for c5")
    ap.add_argument("--orbit-deg", type=float, default=1.0, help="camera orbit per frame in degrees (single-sample workloads; 0 = static camera)")
    ap.add_argument("--no-hbm-leg", action="store_true", help="skip the c5-class (20 M triangles / 16 M incoherent rays) leg of the default line")
    ap.add_argument("--hbm-mtris", type=float, default=20.0); ap.add_argument("--hbm-mrays", type=float, default=16.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--c5-mtris", type=float, default=100.0, help="workload c5: millions of soup triangles")
    ap.add_argument("--c5-mrays", type=float, default=16.0, help="workload c5: millions of incoherent rays per launch")
    ap.add_argument("--c5-total-mrays", type=float, default=256.0, help="workload c5: millions of rays of the whole batch (split over the ranks)")
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
