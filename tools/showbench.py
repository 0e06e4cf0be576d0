"""One-line summary of bench.py JSON files:  python tools/showbench.py gpurun_out/x.json ..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        g = d.get("general_path") or {}
        print("%-44s value %9.0f  fwd-only %9.0f  e2e %9.0f  kernel_ms %s  path %s%s" % (
            f.split("/")[-1], d["value"], d["forward_only"]["value"], d["e2e"]["value"],
            {k: round(v, 2) for k, v in d["kernel_ms"].items()}, d["kernel_path"].split()[0],
            ("  general %.0f" % g["value"]) if g else ""))
    except Exception as e:  # noqa: BLE001
        print("%-44s unreadable (%s)" % (f.split("/")[-1], e))
