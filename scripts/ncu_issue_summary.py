#!/usr/bin/env python3
"""Issue-side utilisation of one frame from the CSV of scripts/gpu_issue.sh.

    python scripts/ncu_issue_summary.py gpurun_out/TAG.csv [--json KEY PATHS]

Per kernel (all launches of the frame summed): time, issue-slot utilisation (smsp__issue_active, % of peak), active lanes
per executed warp instruction, their product / 32 = the share of the SM's lane-issue capacity that did useful work, and
DRAM bytes. The frame figure is the time-weighted mean over the kernels. --json prints the dict bench.py reads from
profiles/latest.json ("issue", and "dram_bytes_per_path" when PATHS, the paths of the frame, is given).
"""
import collections, csv, json, re, sys


def short(name):
    name = name.strip()
    if name.startswith("void "):
        name = name[5:]
    m = re.match(r"(k_[a-z_0-9]+)(<[^>]*>)?", name)          # demangled: k_wf_trace<0, 1, 0, 0>(rt::DScene, ...)
    if m:
        return m.group(1) + (m.group(2) or "").replace(" ", "").replace("(bool)", "").replace("(int)", "")
    m = re.match(r"_Z\d+(k_[a-z_0-9]+?)(I(.*))?E?v", name)    # mangled
    if not m:
        return name[:40]
    vals = re.findall(r"L[bi](\d+)E", m.group(3) or "")
    return m.group(1) + ("<" + ",".join(vals) + ">" if vals else "")


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        d = launches.setdefault(r[ii], {"name": short(r[ki])})
        d[r[mi]] = float(r[vi].replace(",", ""))
    return list(launches.values())


def summarise(launches):
    per = collections.OrderedDict()
    for l in launches:
        t = l["gpu__time_duration.sum"] / 1e6   # ns -> ms
        k = per.setdefault(l["name"], {"ms": 0.0, "n": 0, "issue_t": 0.0, "lanes_i": 0.0, "inst": 0.0, "dram": 0.0, "warps_t": 0.0})
        k["ms"] += t; k["n"] += 1
        k["issue_t"] += t * l["smsp__issue_active.avg.pct_of_peak_sustained_active"] / 100.0
        k["warps_t"] += t * l.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0) / 100.0
        inst = l["smsp__inst_executed.sum"]
        k["inst"] += inst
        k["lanes_i"] += inst * l["smsp__thread_inst_executed_per_inst_executed.ratio"]
        k["dram"] += l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"]
    out = collections.OrderedDict()
    tot_ms = sum(k["ms"] for k in per.values())
    frame = 0.0
    for name, k in per.items():
        issue = k["issue_t"] / k["ms"] if k["ms"] else 0.0
        lanes = k["lanes_i"] / k["inst"] if k["inst"] else 0.0
        out[name] = {"launches": k["n"], "ms": round(k["ms"], 3), "time_share": round(k["ms"] / tot_ms, 4), "issue_active": round(issue, 4),
                     "active_lanes": round(lanes, 2), "issue_x_lanes": round(issue * lanes / 32.0, 4), "warps_active": round(k["warps_t"] / k["ms"], 3) if k["ms"] else 0.0,
                     "dram_gb": round(k["dram"] / 1e9, 3), "dram_gbs": round(k["dram"] / 1e9 / (k["ms"] / 1e3), 1) if k["ms"] else 0.0}
        frame += k["ms"] / tot_ms * issue * lanes / 32.0
    return out, tot_ms, frame, sum(k["dram"] for k in per.values())


if __name__ == "__main__":
    path = sys.argv[1]
    per, tot_ms, frame, dram = summarise(load(path))
    if "--json" in sys.argv:
        i = sys.argv.index("--json")
        key = sys.argv[i + 1]
        paths = float(sys.argv[i + 2]) if len(sys.argv) > i + 2 else 0
        d = {"issue": {"frame_issue_x_lanes": round(frame, 4), "frame_ms_under_ncu": round(tot_ms, 3), "source": path.replace("gpurun_out/", "profiles/").replace("/", "_", 1) if False else path,
                       "what": "time-weighted mean over the kernels of one frame of smsp__issue_active x active lanes per warp instruction / 32 (ncu, scripts/gpu_issue.sh)",
                       "kernels": per}}
        if paths:
            d["dram_bytes_per_path"] = dram / paths
        print(json.dumps({key: d}, indent=1))
    else:
        print("%-28s %4s %9s %6s %7s %6s %8s %6s %9s %8s" % ("kernel", "n", "ms", "share", "issue", "lanes", "iss*ln/32", "warps", "DRAM GB", "GB/s"))
        for name, k in per.items():
            print("%-28s %4d %9.3f %6.3f %7.3f %6.2f %8.3f %6.2f %9.3f %8.1f" % (name, k["launches"], k["ms"], k["time_share"], k["issue_active"], k["active_lanes"],
                                                                              k["issue_x_lanes"], k["warps_active"], k["dram_gb"], k["dram_gbs"]))
        print("frame: %.3f ms under ncu, time-weighted issue x lanes / 32 = %.3f, DRAM %.2f GB" % (tot_ms, frame, dram / 1e9))
