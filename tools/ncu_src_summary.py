"""Summarise `ncu -i X.ncu-rep --page source --print-source sass --csv` output: stall samples by reason, by opcode, top lines."""
import collections
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[2:] if len(r) > 2 and r[2].strip().isdigit()]
    tot = sum(int(r[2]) for r in body)
    print(rows[0][1][:120])
    print("total samples", tot, " instructions executed (warp):", sum(int(r[col['Instructions Executed']]) for r in body))
    by_reason = collections.Counter()
    by_op = collections.Counter()
    by_op_exec = collections.Counter()
    for r in body:
        for h in reasons:
            by_reason[h] += int(r[col[h]] or 0)
        t = r[1].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        by_op[op] += int(r[2])
        by_op_exec[op] += int(r[col['Instructions Executed']])
    print("by reason:", [(k.replace("stall_", ""), v) for k, v in by_reason.most_common(10)])
    print("by opcode (samples, executed):", [(k, v, by_op_exec[k]) for k, v in by_op.most_common(14)])
    for r in sorted(body, key=lambda r: -int(r[2]))[:top]:
        rs = sorted(((int(r[col[h]] or 0), h.replace("stall_", "")) for h in reasons), reverse=True)[:2]
        print(f"{int(r[2]):6d} {100*int(r[2])/tot:5.1f}%  exec={r[col['Instructions Executed']]:>8} {rs}  {r[1].strip()[:80]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
