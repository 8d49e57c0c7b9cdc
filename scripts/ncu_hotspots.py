#!/usr/bin/env python3
"""Where do the executed instructions of one profiled kernel go?

    python scripts/ncu_hotspots.py <report.ncu-rep> <library.so> [--kernel SUBSTR] [--top N] [--depth D]

Joins the SASS page of an `ncu --set full --import-source on` report (per-instruction "Instructions Executed",
"Thread Instructions Executed", stall samples) with the line table of the SAME library (`cuobjdump -xelf` +
`nvdisasm -gi`; the library must have been built with -lineinfo) and prints, per source line and per inlined
call chain root, the share of warp instructions, the average number of active lanes and the stall samples.
The library must be the build that was profiled (the script checks the opcode sequence).
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_page(rep, kernel, index=0):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks, cur = [], None
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["rows"].append(dict(zip(cur["hdr"], r)))
    hits = [b for b in blocks if kernel in b["name"]]
    if len(hits) > index:
        return hits[index]
    raise SystemExit("kernel %r not in report (has: %s)" % (kernel, [b["name"][:60] for b in blocks]))


def line_table(lib, mangled_hint):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    sections = {}
    name, chain, pending = None, [], []
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            name = m.group(1)
            sections[name] = []
            chain, pending = [], []
            continue
        if name is None:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            if pending:
                chain, pending = pending, []
            sections[name].append((int(m.group(1), 16), m.group(2).strip(), tuple(chain)))
    return sections


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("lib")
    ap.add_argument("--kernel", default="k_render")
    ap.add_argument("--index", type=int, default=0, help="n-th launch among those matching --kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--depth", type=int, default=1, help="which level of the inline chain names a 'region' (counted from the kernel body)")
    args = ap.parse_args()
    blk = sass_page(args.report, args.kernel, args.index)
    rows = blk["rows"]
    print("kernel:", blk["name"])
    base = int(rows[0]["Address"], 16)
    sections = line_table(args.lib, blk["name"])
    # choose the section whose opcode sequence matches
    ops = [r["Source"].split()[0] if not r["Source"].strip().startswith("@") else r["Source"].split()[1] for r in rows]
    best = None
    for name, ins in sections.items():
        if len(ins) != len(rows):
            continue
        o2 = [i[1].split()[0] if not i[1].startswith("@") else i[1].split()[1] for i in ins]
        same = sum(a == b for a, b in zip(ops, o2))
        if best is None or same > best[0]:
            best = (same, name)
    if best is None or best[0] < 0.98 * len(rows):
        raise SystemExit("no section of %s matches the profiled SASS (%d instructions): rebuild mismatch" % (args.lib, len(rows)))
    ins = sections[best[1]]
    print("section:", best[1], "(%d instructions, %d opcodes identical)" % (len(ins), best[0]))
    stall_keys = [k for k in rows[0] if k.startswith("stall_") and "Not Issued" not in k]
    tot_i = tot_t = tot_s = 0
    by_line = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    by_region = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    for r, (off, text, chain) in zip(rows, ins):
        ie, te, sm = int(r["Instructions Executed"] or 0), int(r["Thread Instructions Executed"] or 0), int(r["# Samples"] or 0)
        tot_i += ie; tot_t += te; tot_s += sm
        inner = chain[0] if chain else ("?", 0)
        # chain is innermost first; the last entry is the kernel body line, region = the entry `depth` above it
        region = chain[max(0, len(chain) - 1 - args.depth)] if chain else ("?", 0)
        for key, tab in ((inner, by_line), (region, by_region)):
            t = tab[key]
            t[0] += ie; t[1] += te; t[2] += sm
            for k in stall_keys:
                v = int(r[k] or 0)
                if v:
                    t[3][k] += v
    print("warp instructions %d, thread instructions %d (%.1f lanes), samples %d" % (tot_i, tot_t, tot_t / max(1, tot_i), tot_s))

    def show(tab, title):
        print("\n== %s" % title)
        print("%7s %7s %6s  %-22s %s" % ("inst%", "samp%", "lanes", "where", "top stalls"))
        for key, t in sorted(tab.items(), key=lambda kv: -kv[1][0])[:args.top]:
            st = ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / max(1, t[2])) for k, v in t[3].most_common(3))
            print("%6.2f%% %6.2f%% %6.1f  %-22s %s" % (100.0 * t[0] / max(1, tot_i), 100.0 * t[2] / max(1, tot_s), t[1] / max(1, t[0]),
                                                    "%s:%d" % key, st))
    show(by_region, "by region (inline chain level %d)" % args.depth)
    show(by_line, "by innermost source line")


if __name__ == "__main__":
    main()
