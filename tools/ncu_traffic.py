"""Regenerates profiles/ncu_traffic.json from an `ncu --set full` capture of the 14x14 ROIAlign launch and the bench line of
the SAME build (the source hash ties them together; bench.py only reports `roofline.traffic` when the hash of the library
it loaded equals the recorded one).   python tools/ncu_traffic.py capture.ncu-rep bench_line.json [out.json]"""
import csv, io, json, subprocess, sys

rep, line = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else "profiles/ncu_traffic.json"
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
hdr, units, vals = raw[0], raw[1], raw[2]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
get = lambda k: float(vals[hdr.index(k)]) * scale[units[hdr.index(k)]]
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
d = json.loads([l for l in open(line).read().splitlines() if l.startswith("{")][-1])
rec = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum of ONE 14x14 PyramidROIAlign launch inside the bench step, from "
                   "`ncu --set full --clock-control none` (tools/gpu/ncu_traffic.sh); written by tools/ncu_traffic.py",
       "source_hash": d["lib_source_hash"], "capture": rep.split("/")[-1], "kernel": vals[hdr.index("Kernel Name")][:80],
       "dram_bytes_read": rd, "dram_bytes_write": wr, "roialign_p14_pipeline_bytes": rd + wr,
       "duration_us_under_ncu": get("gpu__time_duration.sum") if units[hdr.index("gpu__time_duration.sum")] in scale else
                                float(vals[hdr.index("gpu__time_duration.sum")]),
       "algorithmic_bytes_per_launch": d["roofline"]["algorithmic_bytes_per_launch"]}
json.dump(rec, open(out, "w"), indent=1)
print(json.dumps(rec))
