#!/usr/bin/env python
"""bench.py — fused framesets/s of the DepthFusion hot path on N B200s (one process per GPU).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

A "step" is one pass of the hot path (filter -> deproject -> transform -> splat -> resolve; the whole
FusionNode::processSyncedFrames device sequence, fusion_node.cpp:739-784) over one batch of
--framesets synthetic framesets per GPU. Workload = BASELINE.json configs[1]: D455 pair 1280x720.
Rank 0 prints ONE JSON line. torch is used for device memory, streams, events and torch.distributed only.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, "pointcloud-depthfusion_b200"))

WORKLOADS = {
    "C1": dict(width=848, height=480, n_cameras=2, desc="C1: D455 pair 848x480 (standalone CPU-runnable case)"),
    "C2": dict(width=1280, height=720, n_cameras=2, desc="C2: D455 pair 1280x720 fusion_node path (BASELINE configs[1])"),
    "C4": dict(width=1280, height=720, n_cameras=4, desc="C4: batched 4-camera 1280x720 virtual-view fusion"),
    "C5": dict(width=3840, height=2160, n_cameras=2, desc="C5: synthetic 4K depth-pair fusion"),
}
DEFAULT_FRAMESETS = {"C1": 1024, "C2": 512, "C4": 256, "C5": 32}
METRIC = "fused_framesets_per_sec"
UNIT = "framesets/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="C2")
    ap.add_argument("--framesets", type=int, default=0, help="framesets per step per GPU (0 = workload default)")
    ap.add_argument("--group", type=int, default=0, help="framesets per launch group (0 = automatic)")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--profile", default="P0")
    ap.add_argument("--density", default="dense")
    ap.add_argument("--horizontal", action="store_true", help="un-rotated virtual view (fusion_node vertical_image:=false)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="skip the NCCL gather-to-rank-0 leg at N>1")
    ap.add_argument("--post-filter", action="store_true", help="also time the batch with the 3x3 gauss post filter folded in")
    ap.add_argument("--registration", action="store_true",
                    help="also time BASELINE configs[2]: deproject + voxel downsample of an 848x480 depth pair (points/s)")
    ap.add_argument("--ref-gpu", action="store_true", help="also time the reference's own kernels (oracle/_ref) on this GPU")
    return ap.parse_args()


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled from a thread
    every ~2 ms (a step lasts a few ms, nvidia-smi -lms cannot sample faster than ~20 ms), nvidia-smi as fallback."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    REASON_BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                   0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []      # nvidia-smi fallback: (t, csv line)
        self.samples = []    # NVML: (t, sm_mhz, power_w, reason bitmask)
        self.marks = []
        self.sm_max = None
        self._stop = False
        self.mode = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = self.gpu
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if self.gpu < len(ids) and ids[self.gpu].isdigit():
                    index = int(ids[self.gpu])
            self.nvml, self.handle = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.mode = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "nvidia-smi"
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop:
            try:
                self.samples.append((time.perf_counter(), float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)),
                                     n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0, int(reasons_fn(self.handle))))
            except Exception:
                pass
            time.sleep(0.002)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    @property
    def ready(self):
        return bool(self.samples or self.lines)

    def mark(self):
        """timestamps bracketing the timed region: samples outside are reported but not used for the median"""
        self.marks.append(time.perf_counter())

    def stop(self):
        lo, hi = (self.marks[0], self.marks[-1]) if len(self.marks) >= 2 else (0.0, float("inf"))
        if self.mode == "nvml":
            self._stop = True
            self.thread.join(timeout=1)
            inside = [x for x in self.samples if lo <= x[0] <= hi]
            used = inside if inside else self.samples
            mask = 0
            for x in used:
                mask |= x[3]
            return {"sm_mhz": statistics.median([x[1] for x in used]) if used else None, "sm_max_mhz": self.sm_max,
                    "power_w_max": max([x[2] for x in used]) if used else None, "samples": len(self.samples),
                    "samples_in_timed_region": len(inside), "source": "NVML polled every 2 ms",
                    "reasons": sorted(name for bit, name in self.REASON_BITS.items() if mask & bit)}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for (t, ln) in self.lines if lo <= t <= hi]
        used = inside if inside else [ln for (_, ln) in self.lines]
        for ln in used:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "samples_in_timed_region": len(inside),
                "source": "nvidia-smi -lms 20", "reasons": sorted(reasons)}


def run_reference(args, world, rank):
    """--impl reference: the reference's CPU path. The reference has no runnable CPU implementation of the
    fusion, so this is the oracle port (oracle/liboracle_fast.so), one frameset per host thread."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import numpy as np
    import dfb_b200
    import pyoracle

    wl = WORKLOADS[args.workload]
    cfg = dfb_b200.make_config(wl["width"], wl["height"], n_cameras=wl["n_cameras"], profile=args.profile, density=args.density)
    cores = pyoracle.max_threads()
    sample = max(cores, 8)  # framesets per step: a bounded sample of the workload, one per thread
    depth, color = pyoracle.synth_pool(cfg, 0, sample)
    for _ in range(max(args.warmup, 1)):
        pyoracle.fuse_batch(cfg.desc, depth, color, sample, mode=0, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pyoracle.fuse_batch(cfg.desc, depth, color, sample, mode=0, n_threads=cores)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "framesets_per_step": sample, "profile": args.profile, "density": args.density},
        "mpix_per_s": value * cfg.mpix_in,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} framesets per step x {args.steps} steps, oracle (sequential z-buffer semantics), "
                                   f"one frameset per OpenMP thread"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(cfg, args):
    """Oracle port timed on the host cores on a bounded sample (~10 s of wall time)."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import pyoracle

    cores = pyoracle.max_threads()
    sample = max(cores, 8)
    depth, color = pyoracle.synth_pool(cfg, 0, sample)
    pyoracle.fuse_batch(cfg.desc, depth, color, sample, mode=0, n_threads=cores)
    t0 = time.perf_counter()
    passes = 0
    while passes < 2 or (time.perf_counter() - t0 < 8.0 and passes < 64):
        pyoracle.fuse_batch(cfg.desc, depth, color, sample, mode=0, n_threads=cores)
        passes += 1
    dt = time.perf_counter() - t0
    # single-thread figure on the same sample
    t1 = time.perf_counter()
    pyoracle.fuse_batch(cfg.desc, depth, color, min(sample, 4), mode=0, n_threads=1)
    dt1 = time.perf_counter() - t1
    return {"value": sample * passes / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{sample} framesets x {passes} passes of the same workload, oracle/liboracle_fast.so, one frameset per thread",
            "single_thread_value": min(sample, 4) / dt1}


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, world, rank)
        return

    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist
    import dfb_b200
    from dfb_b200 import capi

    assert torch.cuda.is_available(), "bench.py needs a B200: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dfb_b200.init(local_rank)
    L = dfb_b200.lib()

    wl = WORKLOADS[args.workload]
    F = args.framesets or DEFAULT_FRAMESETS[args.workload]
    cfg = dfb_b200.make_config(wl["width"], wl["height"], n_cameras=wl["n_cameras"], profile=args.profile, density=args.density,
                               vertical=not args.horizontal)
    fuser = dfb_b200.Fuser(cfg.desc)
    if args.group:
        fuser.set_group(args.group)
    if args.variant:
        fuser.set_variant(args.variant)

    # ---- synthetic inputs, generated on the device; each rank owns framesets [rank*F, (rank+1)*F) ----------
    stream = torch.cuda.current_stream()
    sh = stream.cuda_stream
    depth = torch.empty((F, cfg.n_cameras, cfg.height, cfg.width), dtype=torch.int16, device="cuda")
    color = torch.empty((F, cfg.n_cameras, cfg.height, cfg.width, 3), dtype=torch.uint8, device="cuda")
    capi.check(L.dfb_synth_generate(cfg.scenes(rank * F, F), F, cfg.cams_array(), cfg.n_cameras, depth.data_ptr(), color.data_ptr(), sh))
    out = torch.empty((F, cfg.out_height, cfg.out_width, 3), dtype=torch.uint8, device="cuda")
    in_bytes = depth.numel() * 2 + color.numel()
    out_bytes = out.numel()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        fuser.fuse_device(depth.data_ptr(), color.data_ptr(), out.data_ptr(), F, sh)

    # ---- device-resident throughput ("value") -------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        for _ in range(20):  # let nvidia-smi start sampling, keep the GPU under the same load meanwhile
            step()
            torch.cuda.synchronize()
            if sampler.ready:
                break
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = dfb_b200.launch_count()
    barrier()
    sampler.mark()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    sampler.mark()
    launches = dfb_b200.launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop() if rank == 0 else None
    value = F * world * args.steps / (ms_total * 1e-3)

    # ---- per-kernel attribution (events around each launch; separate pass, not the timed region) -------------
    prof = fuser.profile_device(depth.data_ptr(), color.data_ptr(), out.data_ptr(), F, sh)
    prof = fuser.profile_device(depth.data_ptr(), color.data_ptr(), out.data_ptr(), F, sh)
    peak, peak_src = peaks()
    step_ms = ms_total / args.steps
    # measured DRAM traffic of the dominant kernel (ncu --set full capture summarised in profiles/), per launch
    traffic = None
    try:
        with open(os.path.join(REPO, "profiles", "r01_traffic.json")) as fh:
            tr = json.load(fh).get(args.workload)
        if tr and prof["splat_launches"] > 0:
            traffic = tr["dram_bytes_per_frameset"] * F / prof["splat_launches"]
    except (OSError, ValueError, KeyError):
        traffic = None
    achieved = cfg.algorithmic_bytes * F / (step_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "traffic_source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/r01_traffic.json" if traffic else None,
        "peak_source": peak_src,
        "kernel": "fused path = splat (filter+deproject+transform+project+atomicMin) + resolve, timed live over the "
                  "timed region on the launching stream",
        "algorithmic_bytes_per_frameset": cfg.algorithmic_bytes,
        "launch_pairs_per_step": prof["splat_launches"],
        "splat_ms_per_step": prof["splat_ms"], "resolve_ms_per_step": prof["resolve_ms"],
        "splat_share": prof["splat_ms"] / max(prof["splat_ms"] + prof["resolve_ms"], 1e-9),
    }

    # ---- same steps with Frameset::filterColor (3x3 gauss, fusion_node.cpp:789) folded in: the next row of the scope table ----
    filtered = None
    if args.post_filter:
        fuser.set_post_filter(1)
        for _ in range(2):
            step()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(args.steps):
            step()
        f1.record(stream)
        barrier()
        fms = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(fms, op=dist.ReduceOp.MAX)
        filtered = {"value": F * world * args.steps / (float(fms.item()) * 1e-3), "unit": UNIT,
                    "note": "splat + resolve + 3x3 gauss post filter (NPP-exact), device-resident"}
        fuser.set_post_filter(0)
        step()  # restore the unfiltered frames in `out`
        torch.cuda.synchronize()

    # ---- end to end through the C ABI with HOST buffers (H2D + fuse + D2H inside the timed region) -----------
    e2e = None
    if not args.no_e2e:
        h_depth = torch.empty(depth.shape, dtype=torch.int16).pin_memory()
        h_color = torch.empty(color.shape, dtype=torch.uint8).pin_memory()
        h_out = torch.empty(out.shape, dtype=torch.uint8).pin_memory()
        h_depth.copy_(depth)
        h_color.copy_(color)
        e2e_steps = max(3, min(args.steps, 5))
        for _ in range(2):
            fuser.fuse_host(h_depth.data_ptr(), h_color.data_ptr(), h_out.data_ptr(), F)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fuser.fuse_host(h_depth.data_ptr(), h_color.data_ptr(), h_out.data_ptr(), F)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ok = bool(torch.equal(h_out.cuda(), out))
        e2e = {"value": F * world * e2e_steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": in_bytes,
               "d2h_bytes_per_step": out_bytes, "steps": e2e_steps, "matches_device_path": ok,
               "timing": "host wall clock around the synchronous dfb_fuser_fuse_host call (pinned host pools), max over ranks"}

    # ---- optional: gather fused frames to rank 0 over NCCL (the only exchange step of the path) -------------
    gather = None
    if world > 1 and not args.no_gather:
        from dfb_b200.sharding import FrameGather

        bufs = [torch.empty_like(out), torch.empty_like(out)]
        gathers = [FrameGather(F * world, out.shape[1:], dst=0) for _ in range(2)]
        busy = [False, False]
        for b in range(2):  # untimed: NCCL sets up its peer-to-peer channels lazily on the first send/recv
            fuser.fuse_device(depth.data_ptr(), color.data_ptr(), bufs[b].data_ptr(), F, sh)
            gathers[b].start(bufs[b]).finish()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for k in range(args.steps):
            b = k & 1
            if busy[b]:
                gathers[b].finish()
            fuser.fuse_device(depth.data_ptr(), color.data_ptr(), bufs[b].data_ptr(), F, sh)
            gathers[b].start(bufs[b])
            busy[b] = True
        for b in range(2):
            if busy[b]:
                gathers[b].finish()
        g1.record(stream)
        barrier()
        gms = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        gather = {"value": F * world * args.steps / (float(gms.item()) * 1e-3), "unit": UNIT,
                  "bytes_into_rank0_per_step": out_bytes * (world - 1),
                  "note": "same steps with every rank's fused frames gathered to rank 0 by NCCL (double-buffered, async)"}

    # ---- reference's own GPU kernels on the same device ("kernel to beat"), optional ---------------------
    ref_gpu = None
    if args.ref_gpu and rank == 0:
        sys.path.insert(0, os.path.join(REPO, "tests"))
        import ref_kernels

        if ref_kernels.available():
            n_ref = min(F, 32)
            ms_fs = ref_kernels.fuse_bench(cfg.desc, depth.data_ptr(), color.data_ptr(), out.data_ptr(), n_ref, iters=3)
            ref_gpu = {"value": 1e3 / ms_fs, "unit": UNIT, "ms_per_frameset": ms_fs,
                       "note": "unmodified reference kernels.cu (sm_100a) replaying Frameset/Pointcloud call order, device-resident"}
            step()  # restore `out`
            torch.cuda.synchronize()

    # ---- BASELINE configs[2]: registration feed (registration_node.cpp:313-333 + registration.cpp:144-150), optional --------
    registration = None
    if args.registration and rank == 0:
        rcfg = dfb_b200.make_config(848, 480, hole_fraction=0.1)
        rd = torch.empty((1, 2, 480, 848), dtype=torch.int16, device="cuda")
        rc_ = torch.empty((1, 2, 480, 848, 3), dtype=torch.uint8, device="cuda")
        capi.check(L.dfb_synth_generate(rcfg.scenes(0, 1), 1, rcfg.cams_array(), 2, rd.data_ptr(), rc_.data_ptr(), sh))
        ws = dfb_b200.Workspace()
        pts = torch.empty((848 * 480, 4), dtype=torch.float32, device="cuda")
        vox = torch.empty((848 * 480, 4), dtype=torch.float32, device="cuda")
        cnt, nv = C.c_uint32(), C.c_uint32()
        n_pts = n_vox = 0
        reps = 10
        for it in range(reps + 2):
            if it == 2:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                n_pts = n_vox = 0
            for c in range(2):
                capi.check(L.dfb_registration_cloud(ws.handle, C.byref(rcfg.desc.cam[c]), rd[0, c].data_ptr(), pts.data_ptr(),
                                                    C.byref(cnt), sh))
                capi.check(L.dfb_voxel_downsample(ws.handle, pts.data_ptr(), cnt.value, 0.01, vox.data_ptr(), C.byref(nv), sh))
                n_pts += cnt.value
                n_vox += nv.value
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        registration = {"workload": "C3: deproject + voxel downsample (leaf 0.01 m) of an 848x480 depth pair", "points_per_s": n_pts / dt,
                        "pairs_per_s": reps / dt, "points_per_pair": n_pts // reps, "voxels_per_pair": n_vox // reps,
                        "timing": "host wall clock around the synchronous C-ABI calls (each returns the count)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(cfg, args)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wl["desc"], "framesets_per_step_per_gpu": F, "cameras": cfg.n_cameras,
                       "resolution": [cfg.width, cfg.height], "fused_resolution": [cfg.out_width, cfg.out_height],
                       "profile": args.profile, "density": args.density, "vertical": not args.horizontal, "framesets_per_launch_group": fuser.group,
                       "l2_policy": f"inputs {in_bytes / 1e6:.0f} MB per step, larger than the 126 MB L2 (no flush needed)",
                       "sharding": "independent framesets per rank, no data-path collective"},
            "mpix_per_s": value * cfg.mpix_in,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if gather is not None:
            line["gather"] = gather
        if registration is not None:
            line["registration"] = registration
        if filtered is not None:
            line["with_post_filter"] = filtered
        if ref_gpu is not None:
            line["reference_gpu_kernels"] = ref_gpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
