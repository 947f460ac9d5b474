"""Summarise `ncu --page source --csv` of one kernel: executed instructions by opcode, the hottest
SASS lines, and stall samples by reason.  Usage: python tools/ncu_source_summary.py file.csv [rows]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
per_row = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops, stall, total, samples = collections.Counter(), collections.Counter(), 0, 0
lines = []
for r in rows[2:]:
    if len(r) < len(hdr) - 5: continue
    sass = r[ix["Source"]].strip()
    ex = int(r[ix["Instructions Executed"]] or 0)
    sm = int(r[ix["# Samples"]] or 0)
    op = sass.split()[0] if sass and not sass.startswith("@") else (sass.split()[1] if len(sass.split()) > 1 else sass)
    ops[op.split(".")[0]] += ex
    total += ex; samples += sm
    lines.append((sm, ex, sass))
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stall[h] += int(r[ix[h]] or 0)
print("instructions executed: %d (%.0f per row)" % (total, total / per_row))
for op, c in ops.most_common(22): print("  %-10s %12d  %6.1f per row  %5.1f%%" % (op, c, c / per_row, 100.0 * c / total))
print("stall samples:", ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / max(1, sum(stall.values()))) for k, v in stall.most_common(8)))
print("hottest lines (samples, executed, sass):")
for sm, ex, sass in sorted(lines, reverse=True)[:18]: print("  %6d %10d  %s" % (sm, ex, sass[:90]))
