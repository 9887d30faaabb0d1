"""Turns the raw ncu artefacts of a gpurun call into the tracked summaries under profiles/.

  python tools/make_profile_summary.py <round-tag> <launches.csv> <full.ncu-rep> [bench.json]

* <tag>_launches.md    every kernel of the profiled command with launch count, total and average device
                       time and its SHARE of the step (cold-cache, serialised: compare shares, not absolutes)
* <tag>_kernels.md     per captured kernel of the --set full pass: duration, DRAM bytes, DRAM %, issue %,
                       occupancy limiters, top stall reasons
* pyramid_traffic.json DRAM bytes per image of the captured pyramid kernels (bench.py reads it into roofline.traffic)
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(tag, path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "")
        agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = [f"# {tag}: launch list (ncu --metrics gpu__time_duration.sum --clock-control none)", "",
           "Cold-cache, serialised launches: the SHARE column is what is comparable with the live bench.", "",
           "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{name}` | {len(v)} | {sum(v)/1e3:.1f} | {sum(v)/len(v)/1e3:.1f} | {100*sum(v)/tot:.1f} % |")
    out.append(f"| total | {sum(len(v) for v in agg.values())} | {tot/1e3:.1f} | | 100 % |")
    pyr = sum(sum(v) for k, v in agg.items() if "blur" in k or "resize" in k or "u8_to_f32" in k)
    out += ["", f"Pyramid + DoG kernels (blur_*, resize): {100*pyr/tot:.1f} % of the device time of a pass."]
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")


def kernels(tag, rep, batch):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB read"), ("dram__bytes_write.sum", "MB written"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
            ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
            ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem)")]
    stalls = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
    out = [f"# {tag}: ncu --set full --clock-control none, one device pass of {batch} 1080p frames (exact blur arithmetic)", "",
           "DRAM % is ncu's gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed (of the device's nominal peak, not of the measured copy bandwidth).", ""]
    total_dram = 0.0
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        out.append(f"## `{name}` grid {r[idx['Grid Size']]}")
        for m, label in want:
            if m in idx:
                if m.startswith("dram__bytes"):
                    mb = float(r[idx[m]].replace(",", "")) * scale.get(rows[1][idx[m]], 1e6) / 1e6
                    out.append(f"* {label}: {mb:.1f}")
                else:
                    out.append(f"* {label}: {r[idx[m]]}")
        units = {h: rows[1][idx[h]] for h in ("dram__bytes_read.sum", "dram__bytes_write.sum") if h in idx}
        for h, u in units.items():
            v = float(r[idx[h]].replace(",", ""))
            if "blur" in name or "resize" in name:   # the pyramid + DoG stage only
                total_dram += v * scale.get(u, 1e6)
        ss = sorted(((float(r[idx[h]].replace(",", "") or 0), h) for h in stalls), reverse=True)[:5]
        out.append("* stalls per issued instruction: " + ", ".join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in ss))
        out.append("")
    open(os.path.join(ROOT, "profiles", f"{tag}_kernels.md"), "w").write("\n".join(out) + "\n")
    return total_dram


if __name__ == "__main__":
    tag, lpath, rep = sys.argv[1:4]
    batch = int(sys.argv[5]) if len(sys.argv) > 5 else 16
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launches(tag, lpath)
    dram = kernels(tag, rep, batch)
    n_k = len([1 for _ in open(os.path.join(ROOT, "profiles", f"{tag}_kernels.md")) if _.startswith("## ")])
    json.dump({"dram_bytes_per_image": dram / batch, "kernels_captured": n_k, "batch": batch,
               "note": "dram__bytes_read.sum + dram__bytes_write.sum summed over the captured blur / resize kernels (the pyramid + DoG stage) of one device pass / batch; "
                       "levels below 200 rows (octaves 3-4, 1.3 % of the bytes) run as strip kernels that the capture's launch limit left out"},
              open(os.path.join(ROOT, "profiles", "pyramid_traffic.json"), "w"), indent=1)
    if len(sys.argv) > 4 and os.path.exists(sys.argv[4]):
        d = json.load(open(sys.argv[4]))
        json.dump(d, open(os.path.join(ROOT, "profiles", f"{tag}_bench.json"), "w"), indent=1)
    print("dram bytes/image of captured kernels:", dram / batch)
