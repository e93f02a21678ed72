"""Build profiles/traffic.json (dram bytes per launch of each layer kernel) from the raw CSV exports of tools/profile_round.sh.
usage: python tools/traffic_from_ncu.py gpurun_out/ncu_full_y.raw.csv:y gpurun_out/ncu_full_both.raw.csv:both"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (16 frames of 1080x1920), ncu --set full --clock-control none, "
                "round 2 kernels (profiles/ncu_r2_metrics_*.txt); layer1 = the fused pair (layers 1 + 2 in one launch)"}
def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
for spec in sys.argv[1:]:
    path, key = spec.rsplit(":", 1)
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ir, iw, kn = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    rec = {}
    for r in data:
        name = r[kn]
        tot = int(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]))
        layer = 0 if "l0_kernel" in name else 1 if "chain2" in name else 3 if "<3, 1, 1," in name else 4 if "<5, 1, 2," in name else None
        if layer is not None:
            rec["layer%d" % layer] = tot
    out["sesr_x2_int8_sim_1080p:" + key] = rec
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
