"""profiles/pan_kernel_traffic.json from an ncu summary written by tools/ncu_summary.py (dram__bytes_read.sum + dram__bytes_write.sum
of ONE pan_fast_kernel launch): what bench.py reports as roofline.traffic.   usage: make_traffic_json.py summary.txt rows"""
import json, os, re, sys
path, rows = sys.argv[1], int(sys.argv[2])
unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
vals = {}
for line in open(path):
    m = re.match(r"(dram__bytes_(?:read|write)\.sum)\s+([0-9.,]+)\s+(\w+)", line)
    if m:
        vals[m.group(1)] = float(m.group(2).replace(",", "")) * unit[m.group(3)]
out = {"workload_rows": rows, "rows": rows, "kernel": "oip::panfast::pan_fast_kernel<0>",
       "dram_bytes_read": vals["dram__bytes_read.sum"], "dram_bytes_write": vals["dram__bytes_write.sum"],
       "dram_bytes_per_launch": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"],
       "source": os.path.relpath(path, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))) + " (one `ncu --set full` capture, tools/collect_profiles.sh)"}
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "pan_kernel_traffic.json"), "w"), indent=1)
print(out)
