"""Where the warp-stall samples of a kernel fall:  ncu -i x.ncu-rep --page source --csv --print-source sass > s.csv;
python tools/ncu_hot.py s.csv [top]   -- samples per code region (between backward-branch targets), stall mix of the top regions."""
import csv
import re
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
ins = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    ins.append((int(r[ix["Address"]], 16), r[ix["Source"]].strip(), int(r[ix["# Samples"]] or 0), int(r[ix["Instructions Executed"]] or 0),
                {h: int(r[ix[h]] or 0) for h in stall_cols}))
base = ins[0][0]
addr_index = {a: i for i, (a, *_rest) in enumerate(ins)}
# loops = backward branches
loops = []
for i, (a, src, *_r) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)", src)
    if m:
        t = int(m.group(1), 16)
        t = t if t >= base else base + t
        if t < a and t in addr_index and i - addr_index[t] < 1200:
            loops.append((addr_index[t], i))
total = sum(x[2] for x in ins)
print("total samples", total)
seen = []
for lo, hi in sorted(loops, key=lambda p: -(sum(x[2] for x in ins[p[0]:p[1] + 1]))):
    if any(lo >= l2 and hi <= h2 for l2, h2 in seen):
        continue
    seen.append((lo, hi))
    body = ins[lo:hi + 1]
    s = sum(x[2] for x in body)
    if len(seen) > top:
        break
    st = Counter()
    for x in body:
        st.update(x[4])
    ops = Counter(re.sub(r"^@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0] for x in body)
    fp64 = ops["DFMA"] + ops["DADD"] + ops["DMUL"]
    execs = max(x[3] for x in body)
    print("loop +%#x..+%#x  %d instr (%d float64)  samples %d (%.1f%%)  executions %d" % (body[0][0] - base, body[-1][0] - base, len(body), fp64, s, 100.0 * s / total, execs))
    print("   stalls:", ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100.0 * v / max(1, sum(st.values()))) for k, v in st.most_common(8)))
    worst = sorted(body, key=lambda x: -x[2])[:8]
    for x in worst:
        print("   %5d  %-60s %s" % (x[2], x[1][:60], ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in Counter(x[4]).most_common(3))))
