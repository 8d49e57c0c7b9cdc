#!/usr/bin/env python3
"""profiles/latest.json from the issue captures of scripts/gpu_final.sh:  python scripts/make_latest_json.py gpurun_out/TAG profiles/PREFIX

Copies TAG/cN_issue.csv to profiles/PREFIX_cN_issue_ncu.csv and writes, per config, the DRAM bytes per path and the
issue-side figures bench.py puts into `roofline` (traffic, issue)."""
import json, os, shutil, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_issue_summary as S

src, prefix = sys.argv[1], sys.argv[2]
frames = {"c1": (850 * 480 * 1, "1 spp"), "c2": (1920 * 1080 * 16, "16 of 64 spp"), "c3": (3840 * 2160 * 2, "2 of 16 spp"), "c4": (3840 * 2160 * 4, "4 of 256 spp"), "c5": (7680 * 4320 * 2, "2 of 1024 spp")}
out = {}
for key, (paths, what) in frames.items():
    csvp = os.path.join(src, key + "_issue.csv")
    if not os.path.exists(csvp):
        continue
    dst = "%s_%s_issue_ncu.csv" % (prefix, key)
    shutil.copyfile(csvp, dst)
    per, tot_ms, frame, dram = S.summarise(S.load(csvp))
    out[key] = {
        "dram_bytes_per_path": dram / paths,
        "source": dst,
        "what": "one full frame at %s (%d paths), every render kernel, ncu --metrics gpu__time_duration,smsp__issue_active,smsp__thread_inst_executed_per_inst_executed,"
                "dram__bytes_read/write --clock-control none (scripts/gpu_issue.sh)" % (what, paths),
        "issue": {"frame_issue_x_lanes": round(frame, 4), "frame_ms_under_ncu": round(tot_ms, 3),
                  "what": "time-weighted mean over the kernels of the frame of smsp__issue_active x active lanes per executed warp instruction / 32",
                  "kernels": {k: {f: v[f] for f in ("ms", "time_share", "issue_active", "active_lanes", "issue_x_lanes", "dram_gbs")} for k, v in per.items() if v["time_share"] >= 0.01}},
    }
json.dump(out, open("profiles/latest.json", "w"), indent=1)
print(json.dumps({k: (round(v["dram_bytes_per_path"]), v["issue"]["frame_issue_x_lanes"]) for k, v in out.items()}))
