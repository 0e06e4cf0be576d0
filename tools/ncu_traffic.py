"""Turns `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum` logs into profiles/dram_traffic.json
(the `roofline.traffic` figure bench.py reports).   python tools/ncu_traffic.py <workload label> <csv> [<csv> ...]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    label, files = sys.argv[1], sys.argv[2:]
    path = os.path.join(ROOT, "profiles", "dram_traffic.json")
    tab = json.load(open(path)) if os.path.isfile(path) else {"source": "ncu --metrics dram__bytes_read.sum,"
                                                                          "dram__bytes_write.sum --clock-control none", "launches": []}
    for fn in files:
        rows = [r for r in csv.reader(l for l in open(fn) if l.startswith('"'))]
        hdr = rows[0]
        ik, im, iu, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
        per = {}
        for r in rows[1:]:
            k = r[ik].split("(")[0].replace("void ", "").replace("gpmpc::", "")
            per.setdefault((r[0], k), {})[r[im]] = float(r[iv].replace(",", "")) * UNIT.get(r[iu], 1.0)
        for (lid, k), m in per.items():
            tab["launches"] = [x for x in tab["launches"] if not (x["kernel"] == k and x["workload"] == label)]
            tab["launches"].append({"kernel": k, "workload": label, "dram_read_bytes": m.get("dram__bytes_read.sum"),
                                    "dram_write_bytes": m.get("dram__bytes_write.sum"),
                                    "dram_bytes_per_launch": m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)})
    json.dump(tab, open(path, "w"), indent=1)
    print(json.dumps(tab, indent=1))


if __name__ == "__main__":
    main()
