#!/usr/bin/env python
"""Benchmark of the SLR-SFS frame-synthesis hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one scene's clip per GPU: N_FRAMES (60) frames of
Euler -> forward+backward splat -> normalise at 768x1024, 64 feature channels
(BASELINE.json configs[1]).  With --gpus N there are N scenes per step and every
scene's frames are sharded over the N ranks (weak scaling, configs[3]).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "synthesized frames/sec at 768x1024 (N=60)"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default=None, choices=[None, "scatter", "gather"],
                    help="joint-block algorithm (default: the fastest available)")
    ap.add_argument("--height", type=int, default=768)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--channels", type=int, default=64)
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--motion", default="A", choices=["A", "B", "C"])
    ap.add_argument("--batch", type=int, default=0, help="frames per gather launch (0 = library default)")
    ap.add_argument("--no-pipeline", action="store_true", help="issue plan/expand/gather on one stream")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=2, help="frames in the CPU baseline sample")
    return ap.parse_args()


def profiled_traffic(name, frames_per_call):
    """DRAM bytes (read + write) per call of entry point `name`, from the committed ncu --set full
    capture (profiles/*/traffic.json holds bytes per frame of its dominant kernel); None if absent."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*", "traffic.json")), reverse=True):
        try:
            with open(path) as fh:
                entry = json.load(fh).get(name)
            if entry:
                return entry["dram_bytes_per_frame"] * frames_per_call
        except Exception:
            continue
    return None


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


# --------------------------------------------------------------------------
# clocks: sampled (NVML) while the timed region runs
# --------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons of one GPU while the timed region runs.  NVML in a thread
    (a sample every ~10 ms, so that even a 50 ms region is seen); `nvidia-smi -lms` as fallback.
    Samples between mark_begin() and mark_end() are the ones reported; if the region was too short
    to catch one, the samples of the whole bracket (warm-up included) are used and that is said."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, cuda_index):
        self.cuda_index = cuda_index
        self.samples = []            # (host time, sm MHz, set of reasons)
        self.smax = None
        self.t0 = self.t1 = None
        self.how = None
        self._stop = threading.Event()
        self._thread = None
        self._proc = None

    # ---- NVML
    def _nvml_handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(self.cuda_index).uuid)
            if not uuid.startswith("GPU-"):
                uuid = "GPU-" + uuid
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except TypeError:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip() != ""]
            idx = int(ids[self.cuda_index]) if ids and all(v.strip().isdigit() for v in ids) else self.cuda_index
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _nvml_loop(self, pynvml, h):
        reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                bits = int(reasons_fn(h))
                self.samples.append((time.perf_counter(), mhz, {n for b, n in self.REASONS.items() if bits & b}))
            except Exception:
                pass
            self._stop.wait(0.01)

    # ---- nvidia-smi fallback
    def _smi_loop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self._proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                mhz = float(parts[1])
                self.smax = float(parts[2])
            except ValueError:
                continue
            self.samples.append((time.perf_counter(), mhz,
                                 {n for n, v in zip(names, parts[4:8]) if v.lower().startswith("active")}))

    def start(self):
        try:
            pynvml, h = self._nvml_handle()
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._nvml_loop, args=(pynvml, h), daemon=True)
            self.how = "nvml"
        except Exception:
            try:
                self._proc = subprocess.Popen(
                    ["nvidia-smi", "-i", str(self.cuda_index), "--query-gpu=" + self.Q,
                     "--format=csv,noheader,nounits", "-lms", "20"],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self._thread = threading.Thread(target=self._smi_loop, daemon=True)
                self.how = "nvidia-smi"
            except Exception:
                self._thread = None
        if self._thread is not None:
            self._thread.start()

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self._thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0,
                    "source": "unavailable (neither NVML nor nvidia-smi)"}
        self._stop.set()
        if self._proc is not None:
            time.sleep(0.05)
            self._proc.terminate()
            try:
                self._proc.wait(timeout=2)
            except Exception:
                self._proc.kill()
        self._thread.join(timeout=2)
        inside = [s for s in self.samples if self.t0 is not None and self.t1 is not None and self.t0 <= s[0] <= self.t1]
        window = "timed region"
        if not inside:
            inside, window = list(self.samples), "warm-up + timed region (timed region too short for a sample)"
        sm = sorted(s[1] for s in inside)
        reasons = sorted(set().union(*[s[2] for s in inside])) if inside else []
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.smax, "reasons": reasons,
                "samples": len(sm), "window": window, "source": self.how}


# --------------------------------------------------------------------------
# the reference arm / cpu baseline: the reference's own kernels on the host cores
# --------------------------------------------------------------------------
def cpu_frames_per_second(args, n_frames, threads=None):
    """Times the CPU path on `n_frames` frames of the workload: Euler (oracle port) +
    the reference's own splat kernels compiled for the CPU (oracle/_ref, all host
    threads) + the reference's torch glue restated in numpy.  Returns (fps, info)."""
    import numpy as np
    import oracle
    from slr_sfs_b200 import workloads
    feat, Z, motion = workloads.scene(args.height, args.width, args.channels, args.motion, seed=0)
    feat, Z, motion = feat.numpy(), Z.numpy(), motion.numpy()
    use_ref = oracle.ref_available()
    threads = threads or os.cpu_count() or 1
    if use_ref:
        splat = lambda x, f: oracle.ref_softsplat_sum(x, f, threads=threads)
        kind = "reference"
    else:
        splat = oracle.softsplat_sum
        kind, threads = "port", 1
    N = args.frames
    picks = [int(round(i * (N - 1) / max(1, n_frames - 1))) for i in range(n_frames)] if n_frames > 1 else [N // 2]
    t0 = time.perf_counter()
    for t in picks:
        oracle.joint_splat_baseline(feat, Z, motion, (0, t, N - 1), splat=splat)
    dt = time.perf_counter() - t0
    info = {"kind": kind, "cores": threads,
            "sample": "%d of %d frames (t=%s) of the %dx%dx%d workload; splat = %s, Euler = oracle C port (1 thread), "
                      "glue = numpy" % (n_frames, N, picks, args.height, args.width, args.channels,
                                        "reference kernel text compiled for CPU (oracle/_ref, OpenMP)" if use_ref
                                        else "oracle C port")}
    return n_frames / dt, info


def reference_gpu_frames_per_second(args, scene, dev):
    """Informational (not the reference arm): the reference's own CUDA kernel on this GPU
    (oracle/_ref/libref_softsplat_gpu.so: kernel text templated by the reference's cupy_kernel())
    driven like forward_flow -- eager torch Euler loop, torch glue, two launches per frame --
    on a 3-frame sample.  None when the library was not prebuilt or the shape is not baked in."""
    try:
        import torch
        from oracle import refgpu
        feat, Z, motion = scene
        if not refgpu.available() or (1, feat.shape[1] + 1, feat.shape[2], feat.shape[3]) not in refgpu.baked_shapes():
            return None
        N = args.frames
        picks = [0, N // 2, N - 1]
        refgpu.reference_frame(feat, Z, motion, (0, 1, N - 1))          # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in picks:
            refgpu.reference_frame(feat, Z, motion, (0, t, N - 1))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        # kernel-only time of the two splat launches of one frame
        fwd = refgpu.euler_integration(motion, N // 2)
        x = torch.cat([feat, Z], 1).contiguous()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        refgpu.softsplat_sum(x, fwd)
        e0.record()
        for _ in range(4):
            refgpu.softsplat_sum(x, fwd)
        e1.record()
        torch.cuda.synchronize()
        return {"value": len(picks) / dt, "unit": UNIT, "sample": "frames t=%s, host wall clock" % picks,
                "splat_launch_ms": e0.elapsed_time(e1) / 4,
                "what": "reference kernel_Softsplat_updateOutput (unmodified text, reference templating, nvcc sm_100a) + "
                        "eager-torch restatement of euler_integration and the forward_flow glue"}
    except Exception as exc:                                                # informational only
        return {"error": repr(exc)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = max(1, args.cpu_frames)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_frames_per_second(args, 1)
    t0 = time.perf_counter()
    fps_list = []
    for _ in range(args.steps):
        fps, info = cpu_frames_per_second(args, n)
        fps_list.append(fps)
    dt = time.perf_counter() - t0
    fps = args.steps * n / sum(n / f for f in fps_list)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, "cpu"),
        "cpu_baseline": dict(info, value=fps, unit=UNIT),
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(args, algo):
    return {"workload": "configs[1]: %dx%d image, %d-ch encoder features, N=%d frames, forward+backward splat + blend"
                        % (args.height, args.width, args.channels, args.frames),
            "height": args.height, "width": args.width, "channels": args.channels, "frames_per_clip": args.frames,
            "motion": args.motion, "algo": algo,
            "l2": "no flush: one frame's working set (features 201 MB + output 201 MB) exceeds the 126 MB L2",
            "parallelism": "frames of every scene sharded over ranks (one scene per rank per step)"}


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
_JSON_OUT = None


def claim_stdout():
    """stdout must carry exactly ONE JSON line, but libraries chat on it (NCCL prints its version
    banner with printf when NCCL_DEBUG=WARN/VERSION, torchrun children inherit it).  Point fd 1 at
    stderr for the whole run and keep the original stdout for the final line."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    args = parse()
    claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__
    __graft_entry__.build()
    import slr_sfs_b200 as pkg
    from slr_sfs_b200 import workloads, _lib
    from slr_sfs_b200.sharding import frame_block

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    H, W, C, N = args.height, args.width, args.channels, args.frames
    P = H * W
    algo = args.algo or ("gather" if hasattr(pkg.JointSplat, "frame") else "scatter")

    # one scene per rank ("encoded" on that rank); every rank synthesises its frame
    # block of every scene.  Host copies live in pinned memory for the e2e leg.
    scenes_host = []
    for s in range(world):
        feat, Z, motion = workloads.scene(H, W, C, args.motion, seed=s)
        scenes_host.append((feat, Z, motion))
    own = tuple(t.pin_memory() for t in scenes_host[rank])
    lo, hi = frame_block(N, rank, world)

    def make_joint(feat, Z, motion, resident=False):
        # resident inputs are complete: the library may start on them while earlier frames still run
        js = pkg.JointSplat(feat, Z, motion, inputs_event=False if resident else None)
        if args.batch:
            js.batch = args.batch
        js.pipeline = not args.no_pipeline
        return js

    n_mine = hi - lo
    # frames requested per frames() call: several batches, so that the library can overlap the
    # index building of one batch with the gather of the previous one
    batch = args.batch or pkg.JointSplat.batch
    nbuf = min(n_mine, 4 * batch)
    # two output buffers: the consumer of one group of frames (the e2e leg's D2H copy) runs beside
    # the synthesis of the next group
    frame_bufs = [torch.empty(nbuf, C, H, W, dtype=torch.float32, device=dev) for _ in range(2)] if algo == "gather" else None
    buf_free = [None, None]          # events after which a buffer may be overwritten

    def synth_block(js, on_frames=None, chunk=None):
        """Synthesise this rank's frame block of one scene in groups of `chunk` frames;
        on_frames(tensor [k,C,H,W]) consumes each finished group (the decoder's place; the e2e leg
        copies them out) and may return an event that says when the group's buffer is free again."""
        if algo == "gather":
            chunk = min(chunk or nbuf, nbuf)
            js.prepare_clip(0, N - 1, lo, hi - lo)        # Euler chains once for this rank's frame block
            for i, b0 in enumerate(range(lo, hi, chunk)):
                nb = min(chunk, hi - b0)
                slot = i & 1
                if buf_free[slot] is not None:
                    torch.cuda.current_stream().wait_event(buf_free[slot])
                    buf_free[slot] = None
                out = js.frames(0, N - 1, b0, nb, out=frame_bufs[slot][:nb])
                if on_frames is not None:
                    buf_free[slot] = on_frames(out)
        else:
            for t in range(lo, hi):
                out = js.frame_scatter((0, t, N - 1))
                if on_frames is not None:
                    on_frames(out)
        return out

    # resident inputs for the `value` leg
    resident = [tuple(t.to(dev) for t in sc) for sc in scenes_host]
    torch.cuda.synchronize()

    def step_resident(record=None):
        for sc in resident:
            out = synth_block(make_joint(*sc, resident=True))
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step_resident()
    barrier()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # live timing of the dominant kernel only (every bracket costs two event records on the stream)
    dominant = "slr_clip_gather" if algo == "gather" else "slr_joint_scatter"
    _lib.kernel_timing(True, only=[dominant])
    barrier()
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    sampler.mark_end()
    clocks = sampler.stop()
    ktimes = _lib.kernel_timing(False)
    launches = _lib.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    # every entry point on ONE stream, after the timed region: kernel times without the overlap of the
    # two-stream pipeline (what the ncu launch list shows as shares)
    iso_steps = max(1, min(3, args.steps))
    no_pipeline, args.no_pipeline = args.no_pipeline, True
    step_resident()
    _lib.kernel_timing(True)
    for _ in range(iso_steps):
        step_resident()
    ktimes_iso = _lib.kernel_timing(False)
    args.no_pipeline = no_pipeline
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
        tl = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(tl)
        launches = int(tl.item())
    frames_total = world * N * args.steps          # world scenes x N frames per step, over all ranks
    value = frames_total / (ms / 1000.0)

    # ---------------- e2e: host buffers in, host buffers out, every step -----------------
    e2e = None
    if not args.no_e2e:
        n_slots = 3
        ring = [torch.empty(1, C, H, W, dtype=torch.float32).pin_memory() for _ in range(n_slots)]
        copy_stream = torch.cuda.Stream(device=dev)
        in_bytes = sum(t.numel() * 4 for t in own)
        out_bytes = (hi - lo) * world * C * P * 4

        def step_e2e():
            # H2D of this rank's scene, broadcast of every scene from its owner, synthesis of
            # this rank's frame block of every scene, D2H of every synthesised frame.
            mine = tuple(t.to(dev, non_blocking=True) for t in own)
            state = {"k": 0}

            def copy_out(frames):
                done = torch.cuda.Event()
                done.record()
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done)
                    for i in range(frames.shape[0]):
                        ring[state["k"] % n_slots].copy_(frames[i:i + 1], non_blocking=True)
                        state["k"] += 1
                    copied = torch.cuda.Event()
                    copied.record()
                return copied         # the group's buffer must not be overwritten before this

            for s in range(world):
                if world > 1:
                    sc = mine if s == rank else tuple(torch.empty_like(t, device=dev) for t in own)
                    for t in sc:
                        dist.broadcast(t, src=s)
                else:
                    sc = mine
                synth_block(make_joint(*sc), copy_out, chunk=batch)
            copy_stream.synchronize()

        for _ in range(min(args.warmup, 1)):
            step_e2e()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        steps_e2e = max(1, min(args.steps, 2))
        t0.record()
        for _ in range(steps_e2e):
            step_e2e()
        t1.record()
        barrier()
        ms_e = t0.elapsed_time(t1)
        if world > 1:
            tms = torch.tensor([ms_e], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms_e = float(tms.item())
        e2e = {"value": world * N * steps_e2e / (ms_e / 1000.0), "unit": UNIT,
               "h2d_bytes_per_step": in_bytes * world, "d2h_bytes_per_step": out_bytes * world,
               "steps": steps_e2e,
               "note": "per step: pinned-host features+Z+motion -> device (+ NCCL broadcast when N>1), every "
                       "synthesised [C,H,W] fp32 frame -> pinned host ring (3 slots) on a copy stream, beside the synthesis "
                       "of the next group of frames (two device buffers)"}

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        # dominant kernel: the one with the largest total time in the timed region
        roof = None
        if ktimes:
            # entry points issued on the side stream are bracketed by events that include their
            # waits; the dominant kernel is looked for among the ones with algorithmic traffic
            cand = {k: v for k, v in ktimes.items() if _lib.algorithmic_bytes(k, C, P) > 0} or ktimes
            name, (tot_ms, calls) = max(cand.items(), key=lambda kv: kv[1][0])
            frames_per_call = (hi - lo) * world * args.steps / calls if name.startswith("slr_clip") else 1
            alg_bytes = _lib.algorithmic_bytes(name, C, P) * frames_per_call
            avg_s = tot_ms / 1000.0 / calls
            achieved = alg_bytes / avg_s / 1e9
            share = tot_ms / (ms if world == 1 else ms)
            roof = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": profiled_traffic(name, frames_per_call),
                    "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_us": avg_s * 1e6,
                    "share_of_step": share,
                    "measured": "CUDA events around every %s call inside the timed region (two-stream pipeline on: "
                                "the side stream's kernels share the SMs with it)" % name}
            if name in ktimes_iso:
                iso_ms, iso_calls = ktimes_iso[name]
                iso_fpc = (hi - lo) * world * iso_steps / iso_calls if name.startswith("slr_clip") else 1
                iso_s = iso_ms / 1000.0 / iso_calls
                iso_ach = _lib.algorithmic_bytes(name, C, P) * iso_fpc / iso_s / 1e9
                roof["single_stream"] = {"achieved": iso_ach, "frac": iso_ach / peak, "avg_launch_us": iso_s * 1e6,
                                         "steps": iso_steps}
            roof["all_kernels_ms_per_frame"] = {k: v[0] / max(1, (hi - lo) * world * iso_steps)
                                                for k, v in ktimes_iso.items()}
            roof["all_kernels_note"] = "single stream, %d steps after the timed region" % iso_steps
        cpu = None
        ref_gpu = None
        if not args.no_cpu_baseline and world == 1:
            fps, info = cpu_frames_per_second(args, max(1, args.cpu_frames))
            cpu = dict(info, value=fps, unit=UNIT)
            ref_gpu = reference_gpu_frames_per_second(args, resident[0], dev)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, algo), "clocks": clocks, "gpu_launches": launches,
            "e2e": e2e, "roofline": roof, "cpu_baseline": cpu, "reference_gpu": ref_gpu,
            "whole_path_roofline_frac": value / world / (peak * 1e9 / (4.0 * P * (4 * C + 5))),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
