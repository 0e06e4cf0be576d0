"""Float64 instructions per element of the sweeps' hot loops, counted in the SASS of the BUILT library
(rl_gp_mpc/_lib/libgpmpc.so) -> profiles/sass_loop_counts.json, which bench.py uses for the float64 roofline.

    python tools/sass_counts.py            # after `make -C .../csrc`; needs cuobjdump (CUDA toolkit), no GPU

Per kernel the hot loops are the backward branches whose body holds >= 100 DFMA/DADD/DMUL (tools/sass_loops.py); a loop
iteration covers 2 rows x 4 columns (uniform forward sweep) or 2 rows x 8 columns (reverse sweep, general kernel) per
lane.  Loops that read iK (LDG.E.128) belong to the diagonal pairs of the general kernel; of the variants with / without
the residual row shift (far-away rows) the cheaper, common one is reported.  The tensor-core sweeps of the large state dimensions (DMMA.8x8x4 in the loop)
cover 32 rows x 8 columns per round (8 elements per lane); a DMMA counts as the 8 warp-wide DFMAs whose work it does."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "data-efficient-reinforcement-learning-with-probabilistic-model-predictive-control_b200",
                   "rl_gp_mpc", "_lib", "libgpmpc.so")


def loops_of(sass):
    ins = []
    for line in sass.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    index = {a: i for i, (a, _) in enumerate(ins)}
    out = []
    for i, (addr, text) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= addr or tgt not in index or i - index[tgt] > 1400:
            continue
        body = [t for _, t in ins[index[tgt]:i + 1]]
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for t in body)
        plain = sum(v for k, v in ops.items() if k.split(".")[0] in ("DFMA", "DADD", "DMUL"))
        dmma = sum(v for k, v in ops.items() if k.startswith("DMMA"))
        fp64 = plain + 8 * dmma     # a DMMA m8n8k4 = 256 FMAs = the work (and the pipe time) of 8 warp-wide DFMAs
        if fp64 >= 100:
            out.append({"instructions": len(body), "float64": fp64, "other": len(body) - plain - dmma, "dmma": dmma,
                        "ldg128": sum(v for k, v in ops.items() if k.startswith("LDG.E.128")),
                        "lds": sum(v for k, v in ops.items() if k.startswith("LDS")),
                        "shfl": sum(v for k, v in ops.items() if k.startswith("SHFL")),
                        "vimnmx": sum(v for k, v in ops.items() if k.startswith("VIMNMX"))})
    return out


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = sorted(set(re.findall(r"Function : (\S+)", names)))
    table = {}
    for E in range(1, 9):
        entry = {}
        for key, pat, elems in (("uniform_fwd", r"uniform_fwd_kernelILi%dELi(128|256)E" % E, 8),
                                ("uniform_bwd", r"uniform_bwd_kernelILi%dE" % E, 16),
                                ("general_grad", r"rollout_kernelILi%dELb1E" % E, 16),
                                ("general_value", r"rollout_kernelILi%dELb0E" % E, 16)):
            fns = [f for f in funcs if re.search(pat, f)]
            if not fns:
                continue
            fn = sorted(fns)[0]   # (uniform forward: the 128-thread build when it exists)
            sass = subprocess.run(["cuobjdump", "-sass", "-fun", fn, lib], capture_output=True, text=True).stdout
            # a sweep iteration evaluates one exp per element: exactly `elems` clamps (VIMNMX) per iteration
            all_loops = loops_of(sass)
            loops = [l for l in all_loops if l["vimnmx"] == elems and not l["dmma"]]
            mma = [l for l in all_loops if l["dmma"] and l["vimnmx"] == 8]   # tensor-core sweeps: 8 elements per lane and round
            loops4 = [l for l in all_loops if l["vimnmx"] == 2 * elems and l["ldg128"] == 0]   # general kernel: 4 rows per lane
            if key.startswith("uniform"):
                sweep, el = (mma, 8) if mma else (loops, elems)
                best = min(sweep, key=lambda l: l["float64"]) if sweep else None
                if best:
                    entry[key] = {"function": fn, "elements_per_iteration": el, "loop": best, "tile_rows": 32 if mma else 64,
                                  "float64_per_element": best["float64"] / el, "other_per_element": best["other"] / el,
                                  "dmma_per_element": best["dmma"] / el}
            else:
                sweeps = loops
                off = [l for l in sweeps if l["ldg128"] == 0]
                dia = [l for l in sweeps if l["ldg128"] >= 8]
                if off and dia:
                    o, d = min(off, key=lambda l: l["float64"]), min(dia, key=lambda l: l["float64"])
                    entry[key] = {"function": fn, "elements_per_iteration": elems, "loop_off_diagonal": o, "loop_diagonal": d,
                                  "float64_per_element_off_diagonal": o["float64"] / elems,
                                  "float64_per_element_diagonal": d["float64"] / elems,
                                  "other_per_element_off_diagonal": o["other"] / elems,
                                  "other_per_element_diagonal": d["other"] / elems}
                    if loops4:      # off-diagonal pairs when NP is a multiple of 128 (gen_cols4: 2 * elems per iteration)
                        o4 = min(loops4, key=lambda l: l["float64"])
                        entry[key].update({"loop_off_diagonal_rows4": o4,
                                           "float64_per_element_off_diagonal_rows4": o4["float64"] / (2 * elems),
                                           "other_per_element_off_diagonal_rows4": o4["other"] / (2 * elems)})
        if entry:
            table[str(E)] = entry
    out = os.path.join(ROOT, "profiles", "sass_loop_counts.json")
    json.dump({"library": os.path.relpath(lib, ROOT), "how": "tools/sass_counts.py (cuobjdump -sass of the built library)",
               "state_dims": table}, open(out, "w"), indent=1, sort_keys=True)
    for E, e in sorted(table.items()):
        print("E=%s:" % E, {k: {kk: round(vv, 2) for kk, vv in v.items() if kk.startswith(("float64_per", "other_per"))} for k, v in e.items()})


if __name__ == "__main__":
    main()
