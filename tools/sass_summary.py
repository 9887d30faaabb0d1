"""SASS evidence for profiles/: opcode histogram of libsift_gpu.so (whole library and the hot kernels), from cuobjdump -sass.
  python tools/sass_summary.py <tag>   -> profiles/<tag>_sass.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sift_b200", "libsift_gpu.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, per, total = None, collections.OrderedDict(), collections.Counter()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void siftgpu::", "").replace("siftgpu::", "")
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        op = m.group(1)
        per[cur][op] += 1
        total[op] += 1
key = ["UTMALDG.3D", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "LDS.128", "LDS.64", "STS.128", "STG.E.64", "LDG.E.128.CONSTANT", "LDG.E.128"]
out = [f"# {tag}: SASS of sift_b200/libsift_gpu.so (cuobjdump -sass, sm_100a)", "",
       "Whole library, instructions that show how the hot path is built (TMA loads, mbarrier transactions, packed fp32):", "",
       "| opcode | count |", "|---|---:|"]
agg = collections.Counter()
for op, n in total.items():
    for k in key:
        if op == k or op.startswith(k + "."):
            agg[k] += n
for k in key:
    out.append(f"| `{k}` | {agg[k]} |")
out += ["", f"No `UTMASTG` ({sum(n for o, n in total.items() if o.startswith('UTMASTG'))}), no tcgen05 / UTCMMA "
        f"({sum(n for o, n in total.items() if 'UTC' in o and 'MMA' in o)}), no HMMA ({sum(n for o, n in total.items() if o.startswith('HMMA'))}): "
        "no stage of this path is a contraction (north_star).", "",
        "Hot kernels: instructions in the kernel's SASS by class (static counts; the main loop is the bulk of each).", "",
        "| kernel | total | packed fp32 (FFMA2/FMUL2/FADD2) | scalar fp32 | LDS | STS | LDG | STG | UTMALDG | other |", "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
hot = [k for k in per if re.search(r"blur_slide_kernel<(5|7|10|14), (true|false|0|1|\(bool\)[01]), |blur_slide_dec_kernel<(7|10)|extrema_mask_kernel|descriptor_kernel|orientation_kernel|eliminate_kernel", k)]
for k in hot:
    c = per[k]
    tot = sum(c.values())
    cls = lambda pred: sum(n for o, n in c.items() if pred(o))
    packed = cls(lambda o: o.split(".")[0] in ("FFMA2", "FMUL2", "FADD2"))
    scalar = cls(lambda o: o.split(".")[0] in ("FFMA", "FMUL", "FADD", "FMNMX", "FMNMX3", "FSETP", "DFMA", "DMUL", "DADD"))
    lds, sts = cls(lambda o: o.startswith("LDS")), cls(lambda o: o.startswith("STS"))
    ldg, stg = cls(lambda o: o.startswith("LDG") or o.startswith("LD.")), cls(lambda o: o.startswith("STG") or o.startswith("ST."))
    tma = cls(lambda o: o.startswith("UTMALDG"))
    out.append(f"| `{k}` | {tot} | {packed} | {scalar} | {lds} | {sts} | {ldg} | {stg} | {tma} | {tot - packed - scalar - lds - sts - ldg - stg - tma} |")
open(os.path.join(ROOT, "profiles", f"{tag}_sass.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:26]))
