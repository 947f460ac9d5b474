"""Per-phase breakdown of one kernel launch from an `ncu --set full --import-source on` report.

ncu's SASS source page has no line numbers; `nvdisasm -gi` of the same cubin has them, with the
inlining chain.  The two listings hold the same instructions in the same order, so row i of one is
row i of the other.  Every instruction is attributed to the OUTERMOST line of its inlining chain
(the line inside the kernel body) and lines are grouped into phases given on the command line.

  cuobjdump -xelf all approximatenn_b200/libann_b200_f32.so        # -> annb_leaf.sm_100a.cubin
  nvdisasm -gi annb_leaf.sm_100a.cubin > leaf.sass
  ncu -i rep.ncu-rep --page source --csv --print-source sass --launch-skip 7 --launch-count 1 > l7.csv
  python tools/ncu_phase_summary.py l7.csv leaf.sass '_Z18leaf_screen_kernelILi64E' 85000 \
      tables:293-390 pass1:391-493 scan:494-529 pairs:530-576 exact:577-635 rank:636-678 tail:679-700
The number after the symbol prefix divides the totals (e.g. tiles or rows per launch).
"""
import collections, csv, re, sys

src_csv, sass_file, sym, per = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
phases = []
for spec in sys.argv[5:]:
    name, _, rng = spec.partition(":")
    lo, _, hi = rng.partition("-")
    phases.append((name, int(lo), int(hi)))

# 1. outermost source line of every instruction of the function, in order
lines, cur, inside = [], None, False
pending = []
for ln in open(sass_file):
    if ln.startswith(".text."):
        inside = ln.startswith(".text." + sym)
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        pending.append((m.group(1), int(m.group(2)), "inlined at" in m.group(3)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        if pending:
            outer = [p for p in pending if not p[2]]
            cur = (outer[-1] if outer else pending[-1])[1]
            pending = []
        lines.append(cur)

# 2. ncu rows
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":          # ncu prints the table once per view: keep the first
        break
    if len(r) >= len(hdr) - 5:
        body.append(r)
if len(body) != len(lines):
    sys.exit(f"instruction counts differ: ncu {len(body)} vs nvdisasm {len(lines)} — not the same build?")

def num(r, h):
    try:
        return float(r[ix[h]] or 0)
    except (KeyError, ValueError):
        return 0.0

COLS = [("instr", "Instructions Executed"), ("L1 tags (global)", "L1 Tag Requests Global"),
        ("shared wavefronts", "L1 Wavefronts Shared"), ("L2 sectors", "L2 Theoretical Sectors Global"),
        ("samples", "# Samples")]
agg = collections.OrderedDict((p[0], collections.Counter()) for p in phases)
agg["other"] = collections.Counter()
stall = collections.defaultdict(collections.Counter)
for r, line in zip(body, lines):
    ph = "other"
    for name, lo, hi in phases:
        if line is not None and lo <= line <= hi:
            ph = name
            break
    for short, h in COLS:
        agg[ph][short] += num(r, h)
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stall[ph][h[6:]] += num(r, h)
tot = collections.Counter()
print(f"{'phase':8s} " + " ".join(f"{s:>18s}" for s, _ in COLS) + "   top stall reasons")
for ph, c in agg.items():
    tot.update(c)
    s = stall[ph]
    ssum = sum(s.values()) or 1.0
    top = " ".join(f"{k} {100 * v / ssum:.0f}%" for k, v in s.most_common(3))
    print(f"{ph:8s} " + " ".join(f"{c[s_] / per:18.1f}" if s_ != "samples" else f"{c[s_]:18.0f}" for s_, _ in COLS) + "   " + top)
print(f"{'total':8s} " + " ".join(f"{tot[s_] / per:18.1f}" if s_ != "samples" else f"{tot[s_]:18.0f}" for s_, _ in COLS))
