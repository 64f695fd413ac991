"""Extract the DRAM traffic of the profiled launch from an .ncu-rep into profiles/ncu_traffic.json
(bench.py reports it as roofline.traffic).  usage: python tools/ncu_traffic.py rep bunch_steps [out.json]"""
import csv, io, json, subprocess, sys
rep, bunch = sys.argv[1], int(sys.argv[2])
out = sys.argv[3] if len(sys.argv) > 3 else "profiles/ncu_traffic.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, zip(units, vals)))
def to_bytes(k):
    unit, v = d[k]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(v) * scale
res = {"report": rep.split("/")[-1], "kernel": d["Kernel Name"][1] if "Kernel Name" in d else "propagate_persistent",
       "bunch_steps": bunch, "dram_bytes_read": to_bytes("dram__bytes_read.sum"), "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
       "gpu_time_ms_under_ncu": float(d["gpu__time_duration.sum"][1]) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(d["gpu__time_duration.sum"][0], 1)}
res["dram_bytes_total"] = res["dram_bytes_read"] + res["dram_bytes_write"]
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
