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
    ap.add_argument("--no-gather", action="store_true", help="skip the gather-to-rank-0 legs at N>1")
    ap.add_argument("--no-workloads", action="store_true", help="skip the short C1 / C4 / C5 / C3 legs (BASELINE configs 0, 2, 3, 4)")
    ap.add_argument("--no-shim", action="store_true", help="skip the drop-in class API leg and the reference's own GPU kernels")
    ap.add_argument("--post-filter", action="store_true", help="also time the batch with the 3x3 gauss post filter folded in")
    ap.add_argument("--registration", action="store_true", help="(default on) kept for compatibility")
    ap.add_argument("--ref-gpu", action="store_true", help="(default on) kept for compatibility")
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
    # host threads = the cores this process may run on; torch.distributed.run exports OMP_NUM_THREADS=1, which the oracle
    # ignores on purpose (pyoracle.max_threads / dfo_fuse_batch pass the count explicitly)
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
    if cores < 8:
        line["warning"] = f"only {cores} host cores available to this process: the CPU arm is under-provisioned"
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


def parity_spot_check(cfg, out, F):
    """The checker leg (oracle as the CHECKER of the timed batch, never the thing measured): a handful of framesets of
    the batch the timed region just fused — the first ones, both sides of the first key-ring reuse (6 slots), the
    middle and the end — compared bit for bit with the oracle's canonical splat."""
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import pyoracle

    want = [0, 1, 5, 6, 7, F // 2 - 1, F // 2, F - 7, F - 6, F - 1]
    idx = sorted({i for i in want if 0 <= i < F})
    mismatched = 0
    for f in idx:
        d, c = pyoracle.synth_pool(cfg, f, 1)
        exp = pyoracle.fuse_frameset(cfg.desc, list(d[0]), list(c[0]), mode=1)["image"]
        mismatched += int((out[f].cpu().numpy() != exp).any(axis=2).sum())
    return {"framesets": len(idx), "indices": idx, "mismatches": mismatched,
            "what": "fused frames of the timed batch vs oracle (canonical minimum-key splat), bit for bit"}


def timed_steps(fuser, depth, color, out, F, steps, warmup, stream, barrier, world, dist, torch):
    """W warm-up + K timed passes of the device-resident hot path; returns max-over-ranks milliseconds for the K passes"""
    sh = stream.cuda_stream
    for _ in range(warmup):
        fuser.fuse_device(depth.data_ptr(), color.data_ptr(), out.data_ptr(), F, sh)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fuser.fuse_device(depth.data_ptr(), color.data_ptr(), out.data_ptr(), F, sh)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def workload_leg(name, dfb_b200, capi, L, torch, dist, stream, barrier, world, rank, peak, args):
    """One short device-resident leg of another BASELINE config (same launch, same kernel): value + fraction of the
    HBM roofline. C4 is BASELINE configs[3]: 1024 framesets in total, sharded over the ranks (strong scaling); C1 / C5
    keep a fixed batch per GPU (weak)."""
    wl = WORKLOADS[name]
    per_gpu = {"C1": 512, "C4": max(1, 1024 // world), "C5": 16}[name]
    cfg = dfb_b200.make_config(wl["width"], wl["height"], n_cameras=wl["n_cameras"], profile=args.profile, density=args.density)
    fuser = dfb_b200.Fuser(cfg.desc)
    sh = stream.cuda_stream
    depth = torch.empty((per_gpu, cfg.n_cameras, cfg.height, cfg.width), dtype=torch.int16, device="cuda")
    color = torch.empty((per_gpu, cfg.n_cameras, cfg.height, cfg.width, 3), dtype=torch.uint8, device="cuda")
    capi.check(L.dfb_synth_generate(cfg.scenes(rank * per_gpu, per_gpu), per_gpu, cfg.cams_array(), cfg.n_cameras,
                                    depth.data_ptr(), color.data_ptr(), sh))
    out = torch.empty((per_gpu, cfg.out_height, cfg.out_width, 3), dtype=torch.uint8, device="cuda")
    steps = 5
    ms = timed_steps(fuser, depth, color, out, per_gpu, steps, 3, stream, barrier, world, dist, torch)
    value = per_gpu * world * steps / (ms * 1e-3)
    res = {"workload": wl["desc"], "value": value, "unit": UNIT, "framesets_per_step_per_gpu": per_gpu,
           "framesets_per_step_total": per_gpu * world, "scaling": "strong" if name == "C4" else "weak",
           "mpix_per_s": value * cfg.mpix_in,
           "roofline_frac": value / world * cfg.algorithmic_bytes / 1e9 / peak,
           "input_bytes_per_step_per_gpu": depth.numel() * 2 + color.numel()}
    fuser.close()
    del depth, color, out
    torch.cuda.empty_cache()
    return res


def shim_leg(cfg, depth, color, torch, keep_input=None, frames=2000):
    """The drop-in CLASS API (include/pointcloud_processing/*.h) the way fusion_node uses it: the C++ consumer
    tests/cpp/shim_fusion_test.cpp loops FusionNode::processSyncedFrames (fusion_node.cpp:739-789) on one frameset
    pair of this workload, camera images resident on the device — the same conditions under which ref_fuse_bench
    times the reference's own kernels."""
    import struct
    import tempfile

    import numpy as np

    exe = os.path.join(REPO, "pointcloud-depthfusion_b200", "build", "shim_fusion_test")
    if not os.path.exists(exe) or cfg.n_cameras != 2:
        return None
    d = cfg.desc
    intr = d.cam[0].intr
    hdr = struct.pack("<2i4f3f2i2i4f", intr.width, intr.height, intr.fx, intr.fy, intr.ppx, intr.ppy, d.cam[0].depth_scale,
                      d.cam[0].min_depth, d.cam[0].max_depth, d.mirror, 0, d.out_intr.width, d.out_intr.height,
                      d.out_intr.fx, d.out_intr.fy, d.out_intr.ppx, d.out_intr.ppy)
    hdr += struct.pack("<4x") if len(hdr) % 8 else b""
    hdr += np.asarray(cfg.cam_transforms[1], dtype=np.float64).tobytes()
    hdr += np.asarray(cfg.fused_transform, dtype=np.float64).tobytes()
    with tempfile.TemporaryDirectory() as tmp:
        inp, outp = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
        with open(inp, "wb") as f:
            f.write(hdr)
            for c in range(2):
                f.write(depth[0, c].cpu().numpy().tobytes())
                f.write(color[0, c].cpu().numpy().tobytes())
        if keep_input:  # tools/shim_profile.py: the same input file for a run under ncu
            import shutil
            shutil.copyfile(inp, keep_input)
        r = subprocess.run([exe, inp, outp, str(frames)], capture_output=True, text=True, timeout=300)
    if r.returncode != 0:
        return {"error": (r.stdout + r.stderr)[-300:]}
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
    except (ValueError, IndexError):
        return {"error": r.stdout[-300:]}
    return {"value": j["shim_framesets_per_s"], "unit": UNIT, "frames": j["frames"], "matches_first_pass": j["matches_first_pass"],
            "note": "C++ consumer of the drop-in headers looping processSyncedFrames on one frameset pair: Frameset::filterDepth x2 "
                    "(count read back per camera, as the reference does), lazy Pointcloud deproject/transform/append/transform, "
                    "ONE fused launch in project(); camera images resident on the device, host wall clock"}


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
        for _ in range(20):  # let the sampler start, keep the GPU under the same load meanwhile
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

    # ---- the timed batch checked against the oracle (checker leg; rank 0) -----------------------------------
    spot = None
    if rank == 0 and not args.no_cpu_baseline:
        spot = parity_spot_check(cfg, out, F)

    # ---- per-kernel attribution (events around each launch; separate pass, not the timed region) -------------
    prof = fuser.profile_device(depth.data_ptr(), color.data_ptr(), out.data_ptr(), F, sh)
    prof = fuser.profile_device(depth.data_ptr(), color.data_ptr(), out.data_ptr(), F, sh)
    peak, peak_src = peaks()
    step_ms = ms_total / args.steps
    # measured DRAM traffic of the dominant kernel (ncu --set full capture summarised in profiles/), per launch
    traffic, traffic_file = None, None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(REPO, "profiles", name)) as fh:
                tr = json.load(fh).get(args.workload)
            if tr and prof["splat_launches"] > 0:
                traffic, traffic_file = tr["dram_bytes_per_frameset"] * F / prof["splat_launches"], name
                break
        except (OSError, ValueError, KeyError):
            continue
    per_gpu_value = value / world
    achieved = cfg.algorithmic_bytes * per_gpu_value / 1e9
    # the unit that actually binds the kernel (profiles/): packed fp32 instructions of the reference's rounding sequence —
    # 30 (camera without its own transform) / 42 (with) per pixel PAIR, i.e. per pixel as many fp32 lane-operations, on
    # 128 FMA lanes per SM and clock (tools/pipe_probe.cu: FFMA 3.85, FFMA2 1.97 warp-instructions per cycle and SM)
    ops_per_pixel = sum(42.0 if cfg.desc.cam[c].has_transform else 30.0 for c in range(cfg.n_cameras)) / cfg.n_cameras
    sm_mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    fp32_peak = 148 * 128 * sm_mhz * 1e6 / 1e12
    fp32_achieved = per_gpu_value * cfg.mpix_in * 1e6 * ops_per_pixel / 1e12
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
        "traffic_source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, profiles/{traffic_file}" if traffic else None,
        "peak_source": peak_src,
        "kernel": "fuse_stream_kernel = splat (filter+deproject+transform+project+atomicMin) + resolve in ONE persistent launch, "
                  "timed live over the timed region on the launching stream",
        "algorithmic_bytes_per_frameset": cfg.algorithmic_bytes,
        "launches_per_step": prof["splat_launches"],
        "kernel_ms_per_step": prof["splat_ms"] + prof["resolve_ms"],
        "secondary": {"bound": "fp32 pipe (the unit that binds this arithmetic)", "achieved": fp32_achieved, "peak": fp32_peak,
                      "unit": "T lane-op/s", "frac": fp32_achieved / fp32_peak,
                      "fp32_ops_per_pixel": ops_per_pixel,
                      "peak_source": "148 SMs x 128 FMA lanes x SM clock (tools/pipe_probe.cu measures 123 lanes/cycle/SM)"},
    }

    # ---- the other BASELINE configs, short legs of the same launch ---------------------------------------------
    workloads = None
    if not args.no_workloads and args.workload == "C2":
        workloads = {}
        for name in ("C1", "C4", "C5"):
            workloads[name] = workload_leg(name, dfb_b200, capi, L, torch, dist, stream, barrier, world, rank, peak, args)

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
        e2e_value = F * world * e2e_steps / float(dt.item())
        # what limits it: the SAME host buffers copied with nothing else going on, all ranks at once — if the links alone
        # are no faster than the end-to-end run, the end-to-end number is the host <-> device links (shared PCIe switches /
        # host memory), not the library
        d_scratch = torch.empty_like(depth)
        c_scratch = torch.empty_like(color)
        up, down = torch.cuda.Stream(), torch.cuda.Stream()  # both directions of the link at once, like the library's stages
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            with torch.cuda.stream(up):
                d_scratch.copy_(h_depth, non_blocking=True)
                c_scratch.copy_(h_color, non_blocking=True)
            with torch.cuda.stream(down):
                h_out.copy_(out, non_blocking=True)
        torch.cuda.synchronize()
        cdt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(cdt, op=dist.ReduceOp.MAX)
        copy_only = F * world * e2e_steps / float(cdt.item())
        del d_scratch, c_scratch
        e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes,
               "d2h_bytes_per_step": out_bytes, "steps": e2e_steps, "matches_device_path": ok,
               "timing": "host wall clock around the synchronous dfb_fuser_fuse_host call (pinned host pools), max over ranks",
               "link_gbs_per_gpu": (in_bytes + out_bytes) * e2e_steps / float(dt.item()) / 1e9,
               "copies_only": {"value": copy_only, "unit": UNIT,
                               "link_gbs_per_gpu": (in_bytes + out_bytes) * e2e_steps / float(cdt.item()) / 1e9,
                               "note": "the same pinned buffers copied H2D / D2H by all ranks at once with NO kernel: the ceiling "
                                       "the host <-> device links (PCIe switches shared between GPUs, host memory) set for e2e"},
               "limiter": "host<->device links" if copy_only < 1.15 * e2e_value else "library pipeline"}

    # ---- the one exchange step of the path: fused frames to the output rank ---------------------------------
    gather = None
    if world > 1 and not args.no_gather:
        from dfb_b200.sharding import FrameGather, PeerFrameBuffer

        frame_shape = tuple(out.shape[1:])
        # (a) peer stores: every rank's resolve writes its frames straight into rank 0's buffer over NVLink
        peer = [PeerFrameBuffer(F * world, frame_shape, dst=0) for _ in range(2)]
        for b in range(2):
            fuser.fuse_device(depth.data_ptr(), color.data_ptr(), peer[b].slice_ptr(), F, sh)
            torch.cuda.synchronize()
            peer[b].publish()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for k in range(args.steps):
            fuser.fuse_device(depth.data_ptr(), color.data_ptr(), peer[k & 1].slice_ptr(), F, sh)
        g1.record(stream)
        torch.cuda.synchronize()
        peer[0].publish()
        gms = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        peer_value = F * world * args.steps / (float(gms.item()) * 1e-3)
        # KAT B.3-7 on hardware: rank 0 fuses every peer's shard ITSELF and compares with what arrived
        matches = None
        if rank == 0:
            gathered = peer[(args.steps - 1) & 1].frames()
            matches = bool(torch.equal(gathered[:F], out))
            d2 = torch.empty_like(depth)
            c2 = torch.empty_like(color)
            o2 = torch.empty_like(out)
            for r in range(1, world):
                capi.check(L.dfb_synth_generate(cfg.scenes(r * F, F), F, cfg.cams_array(), cfg.n_cameras, d2.data_ptr(), c2.data_ptr(), sh))
                fuser.fuse_device(d2.data_ptr(), c2.data_ptr(), o2.data_ptr(), F, sh)
                torch.cuda.synchronize()
                matches = matches and bool(torch.equal(gathered[r * F:(r + 1) * F], o2))
            del d2, c2, o2
        barrier()
        for pb in peer:
            pb.close()
        # (b) the NCCL gather of round 1 (a second pass over the frames), for comparison
        bufs = [torch.empty_like(out), torch.empty_like(out)]
        gathers = [FrameGather(F * world, out.shape[1:], dst=0) for _ in range(2)]
        busy = [False, False]
        for b in range(2):  # untimed: NCCL sets up its peer-to-peer channels lazily on the first send/recv
            fuser.fuse_device(depth.data_ptr(), color.data_ptr(), bufs[b].data_ptr(), F, sh)
            gathers[b].start(bufs[b]).finish()
        barrier()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record(stream)
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
        n1.record(stream)
        barrier()
        nms = torch.tensor([n0.elapsed_time(n1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(nms, op=dist.ReduceOp.MAX)
        ingress = out_bytes * (world - 1) * args.steps / (float(gms.item()) * 1e-3) / 1e9
        gather = {"value": peer_value, "unit": UNIT, "bytes_into_rank0_per_step": out_bytes * (world - 1),
                  "rank0_ingress_gbs": ingress, "nvlink_ingress_ceiling_gbs": 900.0,
                  "matches_single_gpu": matches,
                  "method": "peer stores: rank 0's frame buffer is IPC-mapped by every rank and passed as d_out, the resolve's "
                            "16-byte stores cross NVLink while the kernel computes (no second pass)",
                  "nccl_gather_value": F * world * args.steps / (float(nms.item()) * 1e-3),
                  "note": "value = the same steps with every fused frame landing in rank 0's memory; nccl_gather_value = round 1's "
                          "dist.gather after the launch, double-buffered"}
        del bufs

    # ---- the drop-in class API and the reference's own GPU kernels on the same device (rank 0) ---------------
    ref_gpu, shim = None, None
    if rank == 0 and not args.no_shim:
        sys.path.insert(0, os.path.join(REPO, "tests"))
        import ref_kernels

        if ref_kernels.available():
            n_ref = min(F, 32)
            ms_fs = ref_kernels.fuse_bench(cfg.desc, depth.data_ptr(), color.data_ptr(), out.data_ptr(), n_ref, iters=3)
            ref_gpu = {"value": 1e3 / ms_fs, "unit": UNIT, "ms_per_frameset": ms_fs,
                       "note": "the kernel to beat: unmodified reference kernels.cu (sm_100a, oracle/_ref) replaying the "
                               "Frameset/Pointcloud call order, camera images resident on the device"}
            step()  # restore `out`
            torch.cuda.synchronize()
        shim = shim_leg(cfg, depth, color, torch)
        if shim and ref_gpu and "value" in shim:
            shim["vs_reference_gpu_kernels"] = shim["value"] / ref_gpu["value"]

    # ---- BASELINE configs[2]: registration feed (registration_node.cpp:313-333 + registration.cpp:144-150) --------
    registration = None
    if rank == 0 and not args.no_workloads:
        rcfg = dfb_b200.make_config(848, 480, hole_fraction=0.1)
        rd = torch.empty((1, 2, 480, 848), dtype=torch.int16, device="cuda")
        rc_ = torch.empty((1, 2, 480, 848, 3), dtype=torch.uint8, device="cuda")
        capi.check(L.dfb_synth_generate(rcfg.scenes(0, 1), 1, rcfg.cams_array(), 2, rd.data_ptr(), rc_.data_ptr(), sh))
        ws = dfb_b200.Workspace()
        pts = torch.empty((848 * 480, 4), dtype=torch.float32, device="cuda")
        vox = torch.empty((848 * 480, 4), dtype=torch.float32, device="cuda")
        cnt, nv = C.c_uint32(), C.c_uint32()
        reps = 10
        registration = {"workload": "C3: deproject + voxel downsample (leaf 0.01 m) of an 848x480 depth pair",
                        "timing": "host wall clock around the synchronous C-ABI calls (each returns the count)"}
        # exact grid (one double-precision centroid per voxel, sorted) and PCL's own semantics (ApproximateVoxelGrid: what
        # Registration::voxelFilter hands to fast_gicp, registration.cpp:144-150)
        for mode, fn in (("exact", L.dfb_voxel_downsample), ("pcl_approximate", L.dfb_voxel_downsample_approximate)):
            n_pts = n_vox = 0
            for it in range(reps + 2):
                if it == 2:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    n_pts = n_vox = 0
                for c in range(2):
                    capi.check(L.dfb_registration_cloud(ws.handle, C.byref(rcfg.desc.cam[c]), rd[0, c].data_ptr(), pts.data_ptr(),
                                                        C.byref(cnt), sh))
                    capi.check(fn(ws.handle, pts.data_ptr(), cnt.value, 0.01, vox.data_ptr(), C.byref(nv), sh))
                    n_pts += cnt.value
                    n_vox += nv.value
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            registration[mode] = {"points_per_s": n_pts / dt, "pairs_per_s": reps / dt, "points_per_pair": n_pts // reps,
                                  "output_points_per_pair": n_vox // reps}
        registration["points_per_s"] = registration["exact"]["points_per_s"]
        registration["pairs_per_s"] = registration["exact"]["pairs_per_s"]

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(cfg, args)

    if rank == 0:
        line = {
            "impl": "ours", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wl["desc"], "framesets_per_step_per_gpu": F, "cameras": cfg.n_cameras,
                       "resolution": [cfg.width, cfg.height], "fused_resolution": [cfg.out_width, cfg.out_height],
                       "profile": args.profile, "density": args.density, "vertical": not args.horizontal, "framesets_per_launch_group": fuser.group,
                       "l2_policy": f"inputs {in_bytes / 1e6:.0f} MB per step, larger than the 126 MB L2 (no flush needed)",
                       "sharding": "independent framesets per rank, no data-path collective"},
            "mpix_per_s": value * cfg.mpix_in,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "parity_spot_check": spot,
        }
        if workloads is not None:
            line["workloads"] = workloads
        if gather is not None:
            line["gather"] = gather
        if registration is not None:
            line["registration"] = registration
        if filtered is not None:
            line["with_post_filter"] = filtered
        if ref_gpu is not None:
            line["reference_gpu_kernels"] = ref_gpu
        if shim is not None:
            line["shim"] = shim
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
