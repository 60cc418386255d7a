"""Turns the ncu outputs of a gpurun call into the small JSON summaries committed under profiles/.

    python tools/summarize_ncu.py launches <launches.csv> <out.json>      # per-launch times of the LAST frame + shares
    python tools/summarize_ncu.py full <raw.csv> <out.json> [kernel ...]  # selected metrics of a --set full capture
                                                                          # (raw.csv = `ncu -i x.ncu-rep --page raw --csv`)
    python tools/summarize_ncu.py source <src.csv> <out.json>             # stall samples (`--page source --csv`)
    python tools/summarize_ncu.py digest <raw.csv> <src.csv> <out.json> <kernel> [keypoint_iterations]
                                                                          # the one-file summary bench.py reads `traffic` from
"""
import csv
import json
import sys

# first launch of a RegisterFrame step with device-resident input: round 2's fused sampler, else round 1's first grid pass
FRAME_FIRST_KERNELS = ("k_sample_fused", "k_grid_claim")
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_membar_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
    "smsp__average_warp_latency_issue_stalled_no_instruction.pct", "sm__icc_requests.sum", "sm__icc_requests_lookup_miss.sum",
]


def rows_of(path):
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.DictReader(lines))


def launches(src, dst):
    rows = [r for r in rows_of(src) if r.get("Metric Name") == "gpu__time_duration.sum"]
    seq = [(r["Kernel Name"].split("(")[0], float(r["Metric Value"]) / 1e3, r["Grid Size"], r["Block Size"]) for r in rows]
    # the last frame = from the last but one... find the last occurrence of the first-of-frame kernel that is followed
    # by a complete frame (the frame sub-sampling launches k_grid_claim twice: frame grid, then keypoint grid)
    first = next(k for k in FRAME_FIRST_KERNELS if any(s[0] == k for s in seq))
    starts = [i for i, s in enumerate(seq) if s[0] == first and (i == 0 or seq[i - 1][0] != "k_grid_emit")]
    begin = starts[-1]
    frame = seq[begin:]
    total = sum(s[1] for s in frame)
    out = {"frame": "last frame of the capture (steady state)", "sum_us": total,
           "launches": [{"kernel": k, "us": us, "share": us / total, "grid": g, "block": b} for k, us, g, b in frame]}
    by_kernel = {}
    for k, us, _, _ in frame:
        by_kernel[k] = by_kernel.get(k, 0.0) + us
    out["share_by_kernel"] = {k: v / total for k, v in sorted(by_kernel.items(), key=lambda kv: -kv[1])}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out["share_by_kernel"], indent=1), "sum_us", total, "launches", len(frame))


def full(src, dst, kernels):
    rows = rows_of(src)
    out = {}
    # the raw page is wide: one row per launch, one column per metric (first data row holds the units)
    units = rows[0] if rows and rows[0].get("ID", "") == "" else {}
    for r in rows:
        name = r.get("Kernel Name", "")
        if not name:
            continue
        short = name.split("(")[0].split("<")[0]
        if kernels and not any(k in short for k in kernels):
            continue
        m = {}
        for key in KEEP:
            if key in r and r[key] != "":
                m[key] = {"value": r[key], "unit": units.get(key, "")}
        out.setdefault(short, []).append({"grid": r.get("Grid Size"), "block": r.get("Block Size"), "metrics": m})
    json.dump(out, open(dst, "w"), indent=1)
    for k, v in out.items():
        print(k, len(v), "launch(es)", {a: b["value"] for a, b in v[0]["metrics"].items()
                                       if a in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum")})


def source(src, dst, top=12):
    """`ncu -i x.ncu-rep --page source --csv` → warp-stall samples per reason and the instructions that collect them."""
    rows = list(csv.reader(open(src, newline="")))
    hdr = next(r for r in rows if "Address" in r and "Source" in r)
    data = rows[rows.index(hdr) + 1:]
    col = {h: i for i, h in enumerate(hdr)}
    reasons = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
    total = sum(int(r[col["# Samples"]]) for r in data)
    by_reason = {h: sum(int(r[col[h]] or 0) for r in data) for h in reasons}
    by_reason = {k: v for k, v in sorted(by_reason.items(), key=lambda kv: -kv[1]) if v}
    hot = sorted(data, key=lambda r: -int(r[col["# Samples"]]))[:top]
    out = {"kernel": rows[0][1] if rows and len(rows[0]) > 1 else "", "samples": total,
           "samples_by_stall_reason": by_reason,
           "share_by_stall_reason": {k: v / total for k, v in by_reason.items()},
           "instructions_executed": sum(int(r[col["Instructions Executed"]] or 0) for r in data),
           "sass_instructions": len(data),
           "hottest_instructions": [
               {"sass": r[col["Source"]].strip(), "offset": r[col["Address"]][-5:], "samples": int(r[col["# Samples"]]),
                "share": int(r[col["# Samples"]]) / total, "executed": int(r[col["Instructions Executed"]] or 0),
                "top_reason": max(reasons, key=lambda h: int(r[col[h]] or 0))} for r in hot]}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out["share_by_stall_reason"], indent=1))
    for h in out["hottest_instructions"]:
        print("%5d %5.1f%% %-18s %s" % (h["samples"], 100 * h["share"], h["top_reason"], h["sass"][:70]))


def digest(raw, src, dst, kernel, kp_iters=None):
    """One JSON per capture: duration, DRAM bytes per launch (= `traffic` of the bench line), instructions (and per
    keypoint-iteration when given), occupancy / issue / cache metrics, stall shares."""
    import os
    import tempfile
    tmp = tempfile.mkdtemp()
    full(raw, os.path.join(tmp, "full.json"), [kernel])
    f = json.load(open(os.path.join(tmp, "full.json")))
    name = next(iter(f))
    m = f[name][0]["metrics"]

    SCALE = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def val(key):   # value in us / bytes / plain number (ncu prints scaled units in the raw page)
        try:
            return float(str(m[key]["value"]).replace(",", "")) * SCALE.get(m[key]["unit"], 1.0)
        except Exception:
            return None
    out = {"kernel": name, "grid": f[name][0]["grid"], "block": f[name][0]["block"],
           "duration_us": val("gpu__time_duration.sum") or 0,
           "dram_bytes_per_launch": (val("dram__bytes_read.sum") or 0) + (val("dram__bytes_write.sum") or 0),
           "metrics": {k: v["value"] + (" " + v["unit"] if v["unit"] else "") for k, v in m.items()}}
    insts = val("smsp__inst_executed.sum") or val("sm__inst_executed.sum")
    if insts:
        out["warp_instructions"] = insts
        if kp_iters:
            out["keypoint_iterations"] = kp_iters
            out["warp_instructions_per_keypoint_iteration"] = insts / kp_iters
    if src and os.path.exists(src):
        source(src, os.path.join(tmp, "src.json"))
        sj = json.load(open(os.path.join(tmp, "src.json")))
        out["stall_share"] = sj["share_by_stall_reason"]
        out["sass_instructions"] = sj["sass_instructions"]
        out["hottest_instructions"] = sj["hottest_instructions"][:8]
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k not in ("metrics", "hottest_instructions")}, indent=1))


def lines(src, dst, top=40):
    """`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` → where the issue slots and the non-barrier stall
    samples of the kernel go, per source file and per source line (inlined code is attributed to its own file)."""
    import collections
    agg = collections.defaultdict(lambda: [0, 0, 0])   # (file, line) → non-barrier samples, instructions, barrier samples
    cur, idx = None, None
    for r in csv.reader(open(src)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            idx = {n: i for i, n in enumerate(r)}
        elif idx and r[0] not in ("", "Function Name") and len(r) > 3 and r[2] == "-":
            a = agg[(cur, int(r[0]))]
            bar = int(r[idx["stall_barrier"]] or 0)
            a[0] += int(r[idx["# Samples"]] or 0) - bar
            a[1] += int(r[idx["Instructions Executed"]] or 0)
            a[2] += bar
    ts, ti, tb = (sum(a[k] for a in agg.values()) for k in range(3))
    files = collections.defaultdict(lambda: [0, 0])
    for (f, _), a in agg.items():
        files[f][0] += a[0]
        files[f][1] += a[1]
    out = {"non_barrier_samples": ts, "barrier_samples": tb, "warp_instructions": ti,
           "by_file": [{"file": f, "sample_share": round(a[0] / max(ts, 1), 4), "instruction_share": round(a[1] / max(ti, 1), 4)}
                       for f, a in sorted(files.items(), key=lambda x: -x[1][0])],
           "by_line": [{"file": f, "line": l, "sample_share": round(a[0] / max(ts, 1), 4),
                        "instruction_share": round(a[1] / max(ti, 1), 4), "instructions": a[1]}
                       for (f, l), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]]}
    json.dump(out, open(dst, "w"), indent=1)
    print("non-barrier samples %d, barrier %d; top files:" % (ts, tb))
    for e in out["by_file"][:8]:
        print("  %-24s samples %5.1f%%  instructions %5.1f%%" % (e["file"], 100 * e["sample_share"], 100 * e["instruction_share"]))


if __name__ == "__main__":
    if sys.argv[1] == "lines":
        lines(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "digest":
        digest(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5], float(sys.argv[6]) if len(sys.argv) > 6 else None)
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "source":
        source(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4:])
