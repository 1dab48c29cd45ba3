#!/usr/bin/env python3
"""bench.py — Mrays/s (path segments/s) of the B200 path-tracing backend on BASELINE.json's configs.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME] [--no-per-config]

A STEP is one frame: one pass of the hot path (Raytracer::render, lib.rs:57-117) over the configured image.  The
headline workload (N = 1 and up) is BASELINE.json configs[1]: Cornell box (scenes.rs:350-414), 800x800, 1000 spp, max
depth 50.

  value      path segments/s with the scene resident in HBM: K frames timed with CUDA events on the launching stream
             (one event pair per frame, L2 flushed between frames), max over ranks.
  e2e        the same metric through the reference-facing call with HOST buffers, per step: the already generated
             World is flattened into a fresh scene (rtwh_world_flatten -> emit calls), uploaded and built (rtw_build),
             rendered and read back — N = 1: one `rtw_render` call with a host frame pointer; N > 1: rtw_render_device
             on every rank + the merge + rank 0's device->host copy.  Scene GENERATION (scenes.rs) is not in it: it is the
             caller's input, like the reference's `Vec<Box<dyn Hittable>>`.
  roofline   the kernel with the LARGEST share of the step.  `achieved` = the bytes that must cross HBM for the
             launch (wavefront state streams; node / primitive records only when the scene does not fit L2) / the
             kernel's CUDA-event time, against the measured HBM copy bandwidth (MEASURED_PEAKS.json).  When that
             fraction is below 0.5 the kernel is not bandwidth bound and `limiter` names what the ncu capture of the
             same kernel shows (issue-slot utilisation, lanes per instruction, top stall).  `traffic` = ncu
             dram__bytes_read.sum + dram__bytes_write.sum per launch (profiles/r02_ncu_summary.json), scaled to this
             run's launch size.
  per_config the other BASELINE.json configs (C1 jumpy-balls, C3 cow, C4 monument, C5 stress) in the same line: value,
             e2e and roofline each, at a stated reduced spp for C4 / C5 — so that every record the driver keeps also
             exercises the LBVH traversal kernel (the headline scene is one leaf).
  cpu_baseline  the oracle running the reference's algorithm (flat list + BvhNode, recursive sample_ray, one pixel per
             task on all host cores) on a bounded sample of the workload.

N > 1 (torchrun): one process per GPU, interleaved 32x32 tiles per rank (rtw_render_params part_rank / part_count),
NCCL reduce(SUM) of the accumulation buffer to rank 0 inside every timed step ("strong" scaling: the frame is fixed).
`ranks` lists every rank's render / merge time of the last timed step (what limits the scaling).
`--gpus N` WITHOUT torchrun measures the single-process product path instead: rtw_render(..., gpus = N), one host thread
per device, peer stores into one frame (include/rtw_cuda.h).
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene, width, height, spp)                                   BASELINE.json config
    "cornell-box": ("cornell-box", 800, 800, 1000),                      # configs[1]
    "jumpy-balls": ("jumpy-balls", 400, 225, 100),                       # configs[0]
    "cow": ("cow-lambert-metal", 1920, 1080, 256),                       # configs[2]
    "monument": ("monument-earth", 3840, 2160, 1024),                    # configs[3]
    "stress": ("stress:1000000:1000000", 3840, 2160, 4096),              # configs[4] (1M spheres + 10M triangles)
}
# the per_config block: (workload, spp rendered there, frames timed).  C4 / C5 at a reduced spp with the same per-pass
# shape (SURVEY.md §8d allows it for sweeps; stated in the block)
PER_CONFIG = [("jumpy-balls", 100, 3), ("cow", 256, 2), ("monument", 32, 2), ("stress", 4, 1)]
SEED = 2024
L2_BYTES = 126 << 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cornell-box", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel (stated in config)")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--pool", type=int, default=0)
    ap.add_argument("--slices", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=20.0, help="target duration of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-per-config", action="store_true")
    ap.add_argument("--per-config", default="", help="comma separated workload[:spp] list replacing the default block")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples = []
        self.proc = None
        self.index = index

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()  # the exact process we started
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        smax = None
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                smax = float(s[1])
            except ValueError:
                continue
            for n, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_summary():
    """Kept ncu numbers of the render kernels (profiles/r02_ncu_summary.json, written by tools/ncu_to_json.py from
    `ncu --set full` captures of this round): DRAM bytes, warp instructions, issue-slot utilisation, lanes per
    instruction per (scene, kernel)."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_summary.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def workload_dims(args, name=None, spp=None):
    scene, w, h, s = WORKLOADS[name or args.workload]
    if name is None:
        return scene, args.width or w, args.height or h, args.spp or s
    return scene, w, h, spp or s


def cpu_baseline(args, scene_name, w, h, max_depth=50):
    """The reference's algorithm on the host cores: oracle, reference structure, recursive integrator."""
    import numpy as np
    import raytracer_weekend_b200 as rtw

    orc = rtw.Backend(os.path.join(ROOT, "oracle", "liboracle.so"), "orc_")  # bench.py's cpu_baseline leg
    orc.fn("render_ex").argtypes = [C.c_void_p, C.POINTER(rtw.Camera), C.POINTER(rtw.RenderParams), C.c_void_p,
                                    C.POINTER(rtw.RenderStats), C.c_int, C.c_int, C.c_int]
    # all host cores of the box, stated explicitly: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # silently make the "all host threads" arm single-threaded
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    with rtw.Scene.from_name(orc, scene_name, w / h, seed=SEED) as so:
        cam = so.cameras[0]
        accum = np.zeros((h, w, 3), np.float32)

        def run(spp_begin, spp_end, part_count=1):
            p = so.params(w, h, spp_end, seed=SEED, max_depth=max_depth, sample_begin=spp_begin, sample_end=spp_end,
                          part_rank=0, part_count=part_count)
            st = rtw.RenderStats()
            t0 = time.perf_counter()
            orc.check(orc.fn("render_ex")(so.h, C.byref(cam), C.byref(p), accum.ctypes.data, C.byref(st), 2, 1, cores), "render_ex")
            return st.segments, time.perf_counter() - t0

        # calibration: one sample per pixel on every 64th 32x32 tile (interleaved over the frame); a denser subset
        # when that is too short to time
        for cal in (64, 8, 1):
            seg, dt = run(0, 1, cal)
            if dt >= 0.3 or cal == 1:
                break
        full_1spp = dt * cal
        if full_1spp > args.cpu_seconds:  # a full-frame sample is too long: every part-th tile of the frame, 1 spp
            part = int(min(4096, max(2, round(full_1spp / args.cpu_seconds))))
            spp = 1
        else:
            part = 1
            spp = max(1, min(4096, int(args.cpu_seconds / max(full_1spp, 1e-9))))
        seg, dt = run(1, 1 + spp, part)
    tiles = "the whole frame" if part == 1 else f"every {part}th 32x32 tile of the frame"
    return {"value": seg / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": f"{scene_name} {w}x{h}, {tiles}, {spp} spp (samples 1..{spp}), max depth {max_depth}, "
                      f"{seg} segments in {dt:.2f} s; oracle = C++ port of the reference (Rust toolchain absent), "
                      "reference structure (flat list + BvhNode), recursive sample_ray, OpenMP dynamic over pixels"}, seg, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    scene_name, w, h, spp = workload_dims(args)
    steps = max(1, args.steps)
    per_step = max(2.0, min(args.cpu_seconds, 120.0 / (steps + args.warmup)))
    args.cpu_seconds = per_step
    segs, dts = 0, 0.0
    base = None
    for i in range(args.warmup + steps):
        base, seg, dt = cpu_baseline(args, scene_name, w, h)
        if i >= args.warmup:
            segs += seg
            dts += dt
    value = segs / dts / 1e6
    base["value"] = value
    line = {"metric": "Mrays/s (path segments/s)", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": dts / steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{scene_name} {w}x{h}, {spp} spp, max depth 50 (CPU arm renders a bounded sample per step)"},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


class Bench:
    """One process per GPU (or one process, N GPUs); measures one workload at a time."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import raytracer_weekend_b200 as rtw
        from raytracer_weekend_b200 import dist as rdist

        self.torch, self.dist, self.rtw, self.rdist, self.args = torch, dist, rtw, rdist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
        # `--gpus N` without torchrun: ONE process drives N devices through rtw_render(..., gpus = N)
        self.single_process_gpus = args.gpus if (self.world == 1 and args.gpus > 1) else 1
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.gpu = rtw.cuda_backend()  # raises if librtw_cuda.so is missing: there is no fallback path
        self.stream = torch.cuda.current_stream(self.dev)
        self.flush = torch.empty(256 << 20, device=self.dev, dtype=torch.uint8)  # > 126 MB L2
        self.peak, self.peak_src = hbm_peak()
        self.ncu = ncu_summary()
        self.sm_count = torch.cuda.get_device_properties(self.dev).multi_processor_count

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def allmax(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, x):
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.int64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return int(t.item())

    # ---- one workload ------------------------------------------------------------------------------------------------
    def measure(self, scene_name, w, h, spp, steps, warmup, want_e2e=True, want_roofline=True, sampler=None):
        torch, rtw, rdist = self.torch, self.rtw, self.rdist
        args = self.args
        import numpy as np

        world_obj = rtw.World(scene_name, w / h, seed=SEED)          # Scene::generate, once (scenes.rs:42-60)
        scene = rtw.Scene.from_world(self.gpu, world_obj, device=self.local_rank)
        cam = scene.cameras[0]
        accum = torch.zeros(h * w * 3, device=self.dev, dtype=torch.float32)
        base_params = scene.params(w, h, spp, seed=SEED, pool_size=args.pool, slices=args.slices)
        out = {}
        last = {}

        if self.single_process_gpus > 1:
            host_frame = np.zeros((h, w, 3), np.float32)
            mp = scene.params(w, h, spp, seed=SEED, pool_size=args.pool, slices=args.slices, gpus=self.single_process_gpus)

            def frame():
                st = rtw.RenderStats()
                self.gpu.check(self.gpu.fn("render")(scene.h, C.byref(cam), C.byref(mp), host_frame.ctypes.data, C.byref(st)), "render")
                return st
        else:
            def render_into(p, buf):
                return scene.render_device(cam, p, buf.data_ptr(), self.stream.cuda_stream)

            def frame():
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                p = rdist.partition(base_params, self.rank, self.world) if self.world > 1 else base_params
                st = render_into(p, accum)
                if self.world > 1:
                    e0.record(self.stream)
                    self.dist.reduce(accum, dst=0, op=self.dist.ReduceOp.SUM)
                    e1.record(self.stream)
                    last["merge_events"] = (e0, e1)
                return st

        for _ in range(warmup):
            frame()
        if sampler:
            sampler.start()
            time.sleep(0.3)
        total_ms, segments, launches, step_ms = 0.0, 0, 0, []
        st = None
        for _ in range(steps):
            self.flush.zero_()  # L2 flush between timed iterations
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record(self.stream)
            st = frame()
            e1.record(self.stream)
            self.barrier()
            wall_ms = (time.perf_counter() - t0) * 1e3
            # one process driving several devices: the work is not on this stream, time the call on the host clock
            ms = self.allmax(wall_ms if self.single_process_gpus > 1 else e0.elapsed_time(e1))
            total_ms += ms
            step_ms.append(ms)
            segments += self.allsum(st.segments) if self.world > 1 else st.segments
            launches += st.launches + 1 + (1 if self.world > 1 else 0)  # + the L2 flush fill (+ the NCCL reduce)
        if sampler:
            sampler.stop()
        out["value"] = segments / (total_ms * 1e-3) / 1e6
        out["ms_per_step"] = total_ms / steps
        out["step_ms"] = step_ms
        out["segments_per_step"] = segments // steps
        out["gpu_launches"] = launches
        out["pool"], out["slices"], out["fused"] = int(st.pool_size), int(st.slices), int(st.fused)
        out["num_prims"] = scene.num_prims
        if self.world > 1:  # what limits the scaling: every rank's render and merge time of the last step
            m0, m1 = last["merge_events"]
            mine = torch.tensor([st.ms_render, m0.elapsed_time(m1), float(st.segments)], device=self.dev, dtype=torch.float64)
            allr = [torch.zeros_like(mine) for _ in range(self.world)]
            self.dist.all_gather(allr, mine)
            out["ranks"] = [{"rank": i, "render_ms": round(float(t[0]), 3), "merge_wait_ms": round(float(t[1]), 3),
                             "segments": int(t[2])} for i, t in enumerate(allr)]

        # ---- e2e: host buffers; flatten + upload + build + render + read-back every step -------------------------------
        if want_e2e:
            host_accum = torch.empty(h * w * 3, dtype=torch.float32).pin_memory() if (self.rank == 0 and self.world > 1) else None
            host_frame = np.zeros((h, w, 3), np.float32)   # the caller's host buffer: exists (and is touched) before the call
            host_frame += 1.0
            e2e_seg, e2e_s, h2d = 0, 0.0, 0
            parts = {"flatten_build_ms": 0.0, "render_call_ms": 0.0}
            n_e2e = min(steps, 3)
            for i in range(1 + n_e2e):
                self.barrier()
                t0 = time.perf_counter()
                s2 = rtw.Scene.from_world(self.gpu, world_obj, device=self.local_rank)   # flatten + H2D + LBVH
                t1 = time.perf_counter()
                p2 = s2.params(w, h, spp, seed=SEED, pool_size=args.pool, slices=args.slices, gpus=self.single_process_gpus)
                if self.world == 1:
                    _, st2 = s2.render(s2.cameras[0], p2, out=host_frame)               # rtw_render, HOST frame pointer
                else:
                    st2 = s2.render_device(s2.cameras[0], rdist.partition(p2, self.rank, self.world), accum.data_ptr(),
                                           self.stream.cuda_stream)
                    self.dist.reduce(accum, dst=0, op=self.dist.ReduceOp.SUM)
                    if self.rank == 0:
                        host_accum.copy_(accum, non_blocking=False)                     # D2H of the merged frame
                self.barrier()
                t2 = time.perf_counter()
                h2d = int(s2.build_stats.device_bytes)
                seg = self.allsum(st2.segments) if self.world > 1 else st2.segments
                s2.close()
                if i > 0:  # the first one warms the allocator
                    e2e_seg += seg
                    e2e_s += self.allmax(t2 - t0)
                    parts["flatten_build_ms"] += (t1 - t0) * 1e3 / n_e2e
                    parts["render_call_ms"] += (t2 - t1) * 1e3 / n_e2e
            out["e2e"] = {"value": e2e_seg / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d,
                          "d2h_bytes_per_step": h * w * 3 * 4, "ms_per_step": e2e_s / n_e2e * 1e3,
                          "call": ("rtwh_world_flatten -> rtw_build -> rtw_render(host frame)" if self.world == 1 else
                                   "rtwh_world_flatten -> rtw_build -> rtw_render_device(part) -> NCCL reduce -> D2H on rank 0"),
                          **{k: round(v, 3) for k, v in parts.items()}}

        # ---- roofline of the dominant kernel (this rank's share) -----------------------------------------------------------
        if want_roofline and self.single_process_gpus == 1:
            out["roofline"] = self.roofline(scene, cam, scene_name, w, h, spp, st, out)
        scene.close()
        world_obj.close()
        return out

    def roofline(self, scene, cam, scene_name, w, h, spp, st, out):
        rtw, rdist, args = self.rtw, self.rdist, self.args
        import torch

        accum = torch.zeros(h * w * 3, device=self.dev, dtype=torch.float32)
        part = (lambda p: rdist.partition(p, self.rank, self.world)) if self.world > 1 else (lambda p: p)
        dev_bytes = int(scene.build_stats.device_bytes)
        resident = dev_bytes < L2_BYTES
        prof = self.ncu.get(scene_name, {})
        if st.fused:
            # one-leaf scene, fused kernel: the whole step is this kernel; its only HBM traffic is the slice sums
            ms = out["ms_per_step"]
            seg = out["segments_per_step"] / (self.world if self.world > 1 else 1)
            items = (w * h // max(self.world, 1)) * int(st.slices)
            hbm_bytes = 16.0 * items + (12.0 * w * h / max(self.world, 1) if st.slices > 1 else 0.0) + dev_bytes
            ach = hbm_bytes / (ms * 1e-3) / 1e9
            k = prof.get("k_mega_flat", {})
            r = {"bound": "hbm", "kernel": "k_mega_flat", "share_of_step": 1.0, "achieved": ach, "peak": self.peak, "unit": "GB/s",
                 "frac": ach / self.peak, "peak_source": self.peak_src, "bytes_per_launch": hbm_bytes,
                 "bytes_per_segment": hbm_bytes / max(seg, 1), "launches": 1, "mean_launch_ms": ms,
                 "traffic": (k["dram_bytes_per_segment"] * seg) if k else None,
                 "note": "fused persistent kernel of a one-leaf scene: paths live in registers, the scene in shared memory; the "
                         "only bytes that must cross HBM are the slice sums (16 B per work item) and one read of the scene, so the "
                         "HBM fraction is ~0 by design and the kernel is bound by instruction issue"}
            if k:
                clock_hz = 1.965e9
                issue_peak = self.sm_count * 4 * clock_hz
                inst_s = k["warp_inst_per_segment"] * seg / (ms * 1e-3)
                r["limiter"] = {"what": "instruction issue", "issue_slot_util_pct": k.get("issue_slot_util_pct"),
                                "lanes_per_inst": k.get("lanes_per_inst"), "warp_inst_per_segment": k["warp_inst_per_segment"],
                                "issue_roofline": {"achieved_ginst_s": inst_s / 1e9, "peak_ginst_s": issue_peak / 1e9,
                                                   "frac": inst_s / issue_peak, "peak": "SMs x 4 schedulers x 1.965 GHz"},
                                "top_stalls": k.get("top_stalls"), "source": k.get("source")}
            return r
        # wavefront: per-kernel CUDA-event times (instrumented single-pool path) and the traversal counters
        p_cnt = part(scene.params(w, h, max(1, spp // 50), seed=SEED, pool_size=args.pool, slices=args.slices,
                                  flags=rtw.RTW_RENDER_COUNT_TRAVERSAL))
        sc = scene.render_device(cam, p_cnt, accum.data_ptr(), self.stream.cuda_stream)
        pairs = sc.node_visits / max(sc.segments, 1)
        prims = sc.prim_tests / max(sc.segments, 1)
        prim_bytes = sc.prim_bytes / max(sc.segments, 1)
        node_bytes = float(sc.node_record_bytes) or 64.0
        p_tim = part(scene.params(w, h, spp, seed=SEED, pool_size=args.pool, slices=args.slices, flags=rtw.RTW_RENDER_TIME_KERNELS))
        stt = scene.render_device(cam, p_tim, accum.data_ptr(), self.stream.cuda_stream)
        seg_per_launch = stt.segments / max(stt.iterations, 1)
        ms_all = max(stt.ms_traverse + stt.ms_shade + stt.ms_sort, 1e-9)
        share_t = stt.ms_traverse / ms_all
        # bytes that must cross HBM per segment:
        #   traverse: ray read 32 B + hit record written 8 B; + node and primitive records when the scene exceeds L2
        #   shade:    6 state streams read (hit 8, origin 16, direction 16, throughput 16, state 16, sum 16 = 88 B) and 4
        #             written on a bounce (64 B); + slot / material / geometry / vertex records when the scene exceeds L2
        scene_t = 0.0 if resident else node_bytes * pairs + prim_bytes
        scene_s = 0.0 if resident else (4 + 8 + 32 + 48 + 64)
        kernels = {
            "k_wave_traverse": {"ms": stt.ms_traverse, "bytes_per_segment": 40.0 + scene_t, "share": share_t},
            "k_wave_shade": {"ms": stt.ms_shade, "bytes_per_segment": 152.0 + scene_s + (4.0 if stt.ray_sort else 0.0),
                             "share": stt.ms_shade / ms_all},
        }
        name = max(kernels, key=lambda k: kernels[k]["share"])
        res = {}
        for kname, kv in kernels.items():
            ach = kv["bytes_per_segment"] * stt.segments / max(kv["ms"] * 1e-3, 1e-12) / 1e9
            k = prof.get(kname, {})
            e = {"bound": "hbm", "kernel": kname if not (kname == "k_wave_traverse" and scene.num_prims <= 32) else "k_wave_traverse_flat",
                 "share_of_step": kv["share"], "achieved": ach, "peak": self.peak, "unit": "GB/s", "frac": ach / self.peak,
                 "bytes_per_segment": kv["bytes_per_segment"], "bytes_per_launch": kv["bytes_per_segment"] * seg_per_launch,
                 "launches": stt.iterations, "mean_launch_ms": kv["ms"] / max(stt.iterations, 1),
                 "traffic": (k["dram_bytes_per_segment"] * seg_per_launch) if k else None}
            if ach / self.peak < 0.5 and k:
                e["limiter"] = {"what": k.get("limiter", "latency / instruction issue"), "issue_slot_util_pct": k.get("issue_slot_util_pct"),
                                "lanes_per_inst": k.get("lanes_per_inst"), "top_stalls": k.get("top_stalls"), "source": k.get("source")}
            res[kname] = e
        r = dict(res[name])
        r["peak_source"] = self.peak_src
        r["pairs_per_segment"], r["prim_tests_per_segment"], r["node_record_bytes"] = pairs, prims, node_bytes
        r["scene_bytes"] = dev_bytes
        if stt.ray_sort:  # rtw_raysort.cuh: the iteration's rays are traced in scene-cell order (hierarchies beyond the caches)
            r["ray_sort"] = {"share_of_step": stt.ms_sort / ms_all, "ms_per_iteration": stt.ms_sort / max(stt.iterations, 1),
                             "kernels": "k_raysort_hist / _scan / _scatter (one 8-bit LSD pass over the slot keys)"}
        r["other_kernel"] = res["k_wave_shade" if name == "k_wave_traverse" else "k_wave_traverse"]
        r["note"] = (("scene (%d B) is L1/L2 resident: node / primitive fetches never reach HBM, only the wavefront state streams do"
                      % dev_bytes) if resident else
                     ("scene (%d B) exceeds L2: node / primitive records are random gathers from HBM; algorithmic bytes per segment "
                      "%.0f (pairs x %d B + primitive records)" % (dev_bytes, node_bytes * pairs + prim_bytes, int(node_bytes))))
        if not resident:
            r["gather_peak"] = {"value": 1290.0, "unit": "GB/s", "source": "profiles/r01_gather_peak.txt (64 B independent gathers "
                                "over 1.4 GB on this pool's B200): the bandwidth a tree walk beyond L2 can reach"}
        return r


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    b = Bench(args)
    scene_name, w, h, spp = workload_dims(args)
    sampler = ClockSampler(b.local_rank) if b.rank == 0 else None
    head = b.measure(scene_name, w, h, spp, args.steps, args.warmup, want_e2e=not args.no_e2e,
                     want_roofline=not args.no_roofline, sampler=sampler)

    per_config = None
    if not args.no_per_config and args.workload == "cornell-box" and not (args.spp or args.width or args.height):
        plan = PER_CONFIG
        if args.per_config:
            plan = []
            for tok in args.per_config.split(","):
                nm, _, sp = tok.partition(":")
                plan.append((nm, int(sp) if sp else WORKLOADS[nm][3], 2))
        per_config = {}
        for nm, sp, frames in plan:
            sn, ww, hh, ss = workload_dims(args, nm, sp)
            t0 = time.perf_counter()
            try:
                r = b.measure(sn, ww, hh, ss, frames, 1, want_e2e=not args.no_e2e, want_roofline=not args.no_roofline)
            except Exception as e:  # a config that cannot run (e.g. out of memory) must not take the headline down
                r = {"error": str(e)[:300]}
            full = WORKLOADS[nm][3]
            r["workload"] = f"{sn} {ww}x{hh}, {ss} spp" + (f" of the config's {full}" if ss != full else "") + ", max depth 50"
            r["bench_seconds"] = round(time.perf_counter() - t0, 1)
            per_config[nm] = r

    base = None
    if b.rank == 0 and b.world == 1 and not args.no_cpu_baseline:
        base, _, _ = cpu_baseline(args, scene_name, w, h)

    if b.rank == 0:
        n = b.world if b.world > 1 else b.single_process_gpus
        line = {"metric": "Mrays/s (path segments/s)", "value": head["value"], "unit": "Mrays/s", "n_gpus": n,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{scene_name} {w}x{h}, {spp} spp, max depth 50", "l2": "flushed between timed steps "
                           "(256 MiB fill)", "parallelism": (f"tiles32x{b.world} (one process per GPU, NCCL reduce)" if b.world > 1 else
                                                             (f"tiles32x{n} (one process, rtw_render gpus={n}, peer stores)" if n > 1 else "single")),
                           "schedule": "fused persistent kernel (one-leaf scene)" if head["fused"] else "wavefront",
                           "pool": head["pool"], "slices": head["slices"], "seed": SEED},
                "segments_per_step": head["segments_per_step"], "step_ms": head["step_ms"],
                "clocks": sampler.summary() if sampler else None, "gpu_launches": head["gpu_launches"]}
        for k in ("e2e", "roofline", "ranks"):
            if k in head:
                line[k] = head[k]
        if per_config:
            line["per_config"] = per_config
        if base:
            line["cpu_baseline"] = base
        print(json.dumps(line))
    if b.world > 1:
        b.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
