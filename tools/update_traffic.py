"""Record the DRAM traffic of the layer-1 dW+Adam kernel from an `ncu --set full` report in profiles/ncu_traffic.json,
keyed by the hash of the sources the library was built from (bench.py reports `roofline.traffic` only while it matches).
usage: python tools/update_traffic.py <report.ncu-rep> <precision> <folds> <D> [summary.txt]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mr_gan_b200 import build

rep, prec, folds, D = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
best = None
for k, r in enumerate(rows[2:]):
    if "k_dw_adam_tc" not in r[ix["Kernel Name"]]:
        continue
    rd = float(r[ix["dram__bytes_read.sum"]].replace(",", "")) * scale[units[ix["dram__bytes_read.sum"]]]
    wr = float(r[ix["dram__bytes_write.sum"]].replace(",", "")) * scale[units[ix["dram__bytes_write.sum"]]]
    if best is None or rd + wr > best[0]:
        best = (rd + wr, rd, wr, r[ix["Grid Size"]], r[ix["Kernel Name"]].split("(")[0], k)
assert best, "no k_dw_adam_tc launch in the report"
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
tj = json.load(open(path)) if os.path.exists(path) else {}
tj["dw1/%s" % prec] = {"kernel": "%s, D layer 1, grid %s" % (best[4], best[3]), "folds": folds, "D": D, "bytes": best[0],
                       "dram_read": best[1], "dram_write": best[2], "launch_in_report": best[5],
                       "source": sys.argv[5] if len(sys.argv) > 5 else os.path.basename(rep), "source_hash": build.source_hash()}
json.dump(tj, open(path, "w"), indent=1)
print(json.dumps(tj["dw1/%s" % prec]))
