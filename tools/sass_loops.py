"""Instruction mix of the float64-heavy loops of a kernel:  cuobjdump -sass -fun <mangled> x.o | python tools/sass_loops.py
Prints, for every backward branch whose body holds >= 100 float64 instructions (the sweeps' hot loops), the body length,
the float64 count and the opcode histogram -- the quick check before spending GPU time on a variant (tools/variants.sh)."""
import collections
import re
import sys

ins = []
for line in sys.stdin:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
index = {a: i for i, (a, _) in enumerate(ins)}
for i, (addr, text) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)", text)
    if not m:
        continue
    tgt = int(m.group(1), 16)
    if tgt >= addr or tgt not in index:
        continue
    body = ins[index[tgt]:i + 1]
    if len(body) > 1000:
        continue
    ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in body)
    fp64 = ops["DFMA"] + ops["DADD"] + ops["DMUL"]
    if fp64 >= 100:
        print("loop %#x-%#x: %d instructions, %d float64, %d other | %s" % (
            tgt, addr, len(body), fp64, len(body) - fp64, ", ".join("%s %d" % kv for kv in ops.most_common(12))))
