#!/usr/bin/env python3
"""Distil `ncu --set full` captures into profiles/r02_ncu_summary.json (what bench.py reads for `roofline.traffic` and
`roofline.limiter`).

    tools/ncu_to_json.py profiles/r02_ncu_summary.json  SCENE:KERNEL:REP[:SEGMENTS] ...

SCENE = bench.py scene name, KERNEL = the key bench.py uses (k_mega_flat, k_wave_traverse, k_wave_shade), REP = .ncu-rep
whose first row matching KERNEL is used, SEGMENTS = path segments that launch processed (default: the launch's
entries = grid-independent pool size must then be given).  Existing entries of the output file are kept."""
import csv
import io
import json
import os
import subprocess
import sys


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[0]
    return [{h[i]: r[i] for i in range(len(h))} for r in rows[2:]]


def num(d, k):
    return float(d[k].replace(",", "")) if d.get(k) not in (None, "") else None


def main():
    path = sys.argv[1]
    data = json.load(open(path)) if os.path.exists(path) else {}
    for spec in sys.argv[2:]:
        parts = spec.split(":")
        scene, kernel, rep = parts[0], parts[1], parts[2]
        # scene names contain ':' (stress:a:b): re-join
        if len(parts) > 4:
            scene = ":".join(parts[:-3]) if parts[-1].isdigit() else ":".join(parts[:-2])
            kernel, rep = (parts[-3], parts[-2]) if parts[-1].isdigit() else (parts[-2], parts[-1])
        segments = float(parts[-1]) if parts[-1].replace(".", "").isdigit() else None
        match = [d for d in rows_of(rep) if kernel.replace("k_wave_traverse", "k_wave_traverse") in d["Kernel Name"]]
        if not match:
            print("no row for", spec)
            continue
        d = match[0]
        dur_ns = num(d, "gpu__time_duration.sum")
        unit_scale = 1.0
        dram = (num(d, "dram__bytes_read.sum") or 0) + (num(d, "dram__bytes_write.sum") or 0)
        # ncu prints these in scaled units (Kbyte / Mbyte / Gbyte): re-read with the unit row
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        rr = list(csv.reader(io.StringIO(out)))
        units = {rr[0][i]: rr[1][i] for i in range(len(rr[0]))}
        def scaled(k):
            v = num(d, k) or 0.0
            u = units.get(k, "")
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
        dram = scaled("dram__bytes_read.sum") + scaled("dram__bytes_write.sum")
        dur_s = dur_ns * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}.get(units.get("gpu__time_duration.sum", "ns"), 1e-9)
        inst = num(d, "smsp__inst_executed.sum")
        stalls = {k.split("stalled_")[1].split("_per")[0]: num(d, k) for k in d if "issue_stalled" in k and k.endswith(".ratio") and "not_issued" not in k}
        top = sorted(stalls.items(), key=lambda kv: -(kv[1] or 0))[:5]
        e = {"duration_s_under_ncu": dur_s, "dram_bytes_per_launch": dram, "warp_inst_per_launch": inst,
             "issue_slot_util_pct": num(d, "sm__inst_issued.avg.pct_of_peak_sustained_active"),
             "lanes_per_inst": num(d, "smsp__thread_inst_executed_per_inst_executed.ratio"),
             "warps_active_pct": num(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
             "registers": num(d, "launch__registers_per_thread"),
             "l1_hit_pct": num(d, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": num(d, "lts__t_sector_hit_rate.pct"),
             "top_stalls": {k: round(v, 2) for k, v in top if v is not None},
             "source": f"profiles/{os.path.basename(rep)} is not kept (64 MiB); distilled by tools/ncu_to_json.py from "
                       f"gpurun_out/{os.path.basename(rep)}: ncu --set full --clock-control none, one launch of {d['Kernel Name'][:60]}"}
        if segments:
            e["segments_per_launch"] = segments
            e["dram_bytes_per_segment"] = dram / segments
            e["warp_inst_per_segment"] = inst / segments
        e["dram_gbs_under_ncu"] = dram / dur_s / 1e9
        data.setdefault(scene, {})[kernel] = e
        print(scene, kernel, json.dumps(e)[:300])
    json.dump(data, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
