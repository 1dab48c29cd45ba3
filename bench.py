#!/usr/bin/env python3
"""bench.py — Mrays/s (path segments/s) of the B200 path-tracing backend on BASELINE.json's config.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

A STEP is one frame: one pass of the hot path (Raytracer::render, lib.rs:57-117) over the
configured image.  Default workload (N = 1 and up) = BASELINE.json configs[1]: Cornell box
(scenes.rs:350-414), 800x800, 1000 spp, max depth 50.

  value      path segments/s with the scene resident in HBM: K frames timed with CUDA events on the
             launching stream (one event pair per frame, L2 flushed between frames), max over ranks.
  e2e        the same metric through the reference-facing call with HOST buffers: every step does
             scene flatten + upload + LBVH build + render + device->host read of the frame.
  roofline   the traversal kernel: algorithmic bytes per launch / its mean CUDA-event duration, over
             the measured HBM copy bandwidth (MEASURED_PEAKS.json).  See DESIGN.md §Measurement.
  cpu_baseline  the oracle running the reference's algorithm (flat list + BvhNode, recursive
             sample_ray, one pixel per task on all host cores) on a bounded sample of the workload.

N > 1 (torchrun): one process per GPU, interleaved 32x32 tiles per rank, NCCL reduce(SUM) of the
accumulation buffer to rank 0 inside every timed step ("strong" scaling: the frame is fixed).
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
SEED = 2024


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


def workload_dims(args):
    scene, w, h, spp = WORKLOADS[args.workload]
    return scene, args.width or w, args.height or h, args.spp or spp


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
    vals, segs, dts = [], 0, 0.0
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


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import raytracer_weekend_b200 as rtw
    from raytracer_weekend_b200 import dist as rdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    gpu = rtw.cuda_backend()  # raises if librtw_cuda.so is missing: there is no fallback path

    scene_name, w, h, spp = workload_dims(args)
    stream = torch.cuda.current_stream(dev)
    accum = torch.zeros(h * w * 3, device=dev, dtype=torch.float32)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)  # > 126 MB L2

    scene = rtw.Scene.from_name(gpu, scene_name, w / h, seed=SEED, device=local_rank)
    cam = scene.cameras[0]
    base_params = scene.params(w, h, spp, seed=SEED, pool_size=args.pool, slices=args.slices)

    def render_into(p, buf):
        return scene.render_device(cam, p, buf.data_ptr(), stream.cuda_stream)

    def frame():
        return rdist.render_frame(render_into, base_params, accum)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        frame()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    total_ms, segments, launches = 0.0, 0, 0
    step_ms = []
    for _ in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        st = frame()
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        seg = torch.tensor([st.segments], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(seg, op=dist.ReduceOp.SUM)
        total_ms += float(ms.item())
        step_ms.append(float(ms.item()))
        segments += int(seg.item())
        launches += st.launches + 1  # + the L2 flush fill
    if sampler:
        sampler.stop()
    value = segments / (total_ms * 1e-3) / 1e6

    # ---- e2e: host buffers, scene flatten + upload + build + render + read-back every step -----------
    e2e = None
    if not args.no_e2e:
        host_accum = torch.empty(h * w * 3, dtype=torch.float32).pin_memory() if rank == 0 else None
        e2e_seg, e2e_s = 0, 0.0
        h2d = 0
        for i in range(1 + min(args.steps, 3)):
            barrier()
            t0 = time.perf_counter()
            s2 = rtw.Scene.from_name(gpu, scene_name, w / h, seed=SEED, device=local_rank)  # flatten + H2D + LBVH
            p2 = s2.params(w, h, spp, seed=SEED, pool_size=args.pool, slices=args.slices)
            st2 = rdist.render_frame(lambda p, buf: s2.render_device(s2.cameras[0], p, buf.data_ptr(), stream.cuda_stream),
                                     p2, accum)
            if rank == 0:
                host_accum.copy_(accum, non_blocking=False)  # D2H of the merged frame
            barrier()
            dt = time.perf_counter() - t0
            h2d = int(s2.build_stats.device_bytes)
            seg = torch.tensor([st2.segments], device=dev, dtype=torch.int64)
            if world > 1:
                dist.all_reduce(seg, op=dist.ReduceOp.SUM)
            s2.close()
            if i > 0:  # first one warms the allocator
                e2e_seg += int(seg.item())
                e2e_s += dt
        e2e = {"value": e2e_seg / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": h * w * 3 * 4, "ms_per_step": e2e_s / min(args.steps, 3) * 1e3}

    # ---- roofline of the traversal kernel (rank 0, single GPU share) -----------------------------------
    roofline = None
    if not args.no_roofline:
        peak, peak_src = hbm_peak()
        p_cnt = scene.params(w, h, max(1, spp // 50), seed=SEED, pool_size=args.pool, slices=args.slices,
                             flags=rtw.RTW_RENDER_COUNT_TRAVERSAL)
        p_cnt = rdist.partition(p_cnt, rank, world) if world > 1 else p_cnt
        sc = scene.render_device(cam, p_cnt, accum.data_ptr(), stream.cuda_stream)
        pairs_per_seg = sc.node_visits / max(sc.segments, 1)
        prims_per_seg = sc.prim_tests / max(sc.segments, 1)
        prim_bytes_per_seg = sc.prim_bytes / max(sc.segments, 1)
        # per segment: 64 B per child-pair fetch (32 B when rtw_build chose compact pairs) + geometry bytes of the
        #            primitive tests + ray read 32 B + hit write 8 B (identity slot mapping: no queue entry)
        node_bytes = float(sc.node_record_bytes) or 64.0
        bytes_per_seg = node_bytes * pairs_per_seg + prim_bytes_per_seg + 32 + 8
        p_tim = scene.params(w, h, spp, seed=SEED, pool_size=args.pool, slices=args.slices, flags=rtw.RTW_RENDER_TIME_KERNELS)
        p_tim = rdist.partition(p_tim, rank, world) if world > 1 else p_tim
        stt = scene.render_device(cam, p_tim, accum.data_ptr(), stream.cuda_stream)
        ach = bytes_per_seg * stt.segments / (stt.ms_traverse * 1e-3) / 1e9
        seg_per_launch = stt.segments / max(stt.iterations, 1)
        # HBM-only variant: the wavefront-state bytes of the kernel alone — what must cross HBM
        # when nodes + primitives are cache resident (ray 32 B read, hit 8 B written)
        hbm_only = 40.0 * stt.segments / (stt.ms_traverse * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as fh:
                tj = json.load(fh).get(scene_name, {}).get("k_wave_traverse")
            if tj:  # ncu dram__bytes_read.sum + dram__bytes_write.sum per segment, scaled to this run's launch size
                traffic = tj["dram_bytes_per_segment"] * seg_per_launch
        resident = pairs_per_seg < 64 and scene.build_stats.device_bytes < (100 << 20)
        roofline = {"bound": "hbm", "kernel": "k_wave_traverse_flat" if scene.num_prims <= 32 else "k_wave_traverse", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                    "bytes_per_launch": bytes_per_seg * seg_per_launch,
                    "bytes_per_segment": bytes_per_seg, "pairs_per_segment": pairs_per_seg, "node_record_bytes": node_bytes,
                    "prim_tests_per_segment": prims_per_seg, "launches": stt.iterations,
                    "mean_launch_ms": stt.ms_traverse / max(stt.iterations, 1),
                    "traverse_share_of_step": stt.ms_traverse / max(stt.ms_traverse + stt.ms_shade, 1e-9),
                    "hbm_only": {"bytes_per_segment": 40.0, "achieved": hbm_only, "frac": hbm_only / peak},
                    "note": ("scene is L1/L2 resident (%d B of nodes + primitives): the algorithmic node/primitive bytes are "
                             "served from cache, so `frac` is a cache-bandwidth figure against the HBM peak and may exceed 1; "
                             "`hbm_only` counts the wavefront-state bytes that do cross HBM (ncu traffic agrees). The kernel is "
                             "issue-bound: see profiles/r01_final_ncu_summary.txt" % scene.build_stats.device_bytes)
                    if resident else "scene exceeds L2: node / primitive fetches are random gathers from HBM, which this GPU "
                    "serves at ~1.3 TB/s (tools/gather_peak.cu, profiles/r01_gather_peak.txt) - 20 % of the copy peak used "
                    "as `peak` here; see DESIGN.md section 6"}
        if not resident:
            roofline["gather_peak"] = {"value": 1290.0, "unit": "GB/s", "source": "profiles/r01_gather_peak.txt (64 B independent "
                                       "gathers over 1.4 GB, measured on this pool's B200)"}

    base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        base, _, _ = cpu_baseline(args, scene_name, w, h)

    if rank == 0:
        line = {"metric": "Mrays/s (path segments/s)", "value": value, "unit": "Mrays/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{scene_name} {w}x{h}, {spp} spp, max depth 50", "l2": "flushed between timed steps "
                           "(256 MiB fill)", "parallelism": f"tiles32x{world}" if world > 1 else "single",
                           "pool": int(st.pool_size), "slices": int(st.slices), "seed": SEED},
                "segments_per_step": segments // args.steps, "step_ms": step_ms,
                "clocks": sampler.summary() if sampler else None, "gpu_launches": launches}
        if e2e:
            line["e2e"] = e2e
        if roofline:
            line["roofline"] = roofline
        if base:
            line["cpu_baseline"] = base
        print(json.dumps(line))
    scene.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
