"""Turns gpurun_out/ evidence into the tracked profiles/ summaries (run after scripts/gpu_round.sh)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
rep = os.path.join(G, "prof_r1_final.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
def val(r, k): return float(r[idx[k]])
launch = []
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    short = name.replace("(int)", "").replace("(bool)", "")
    mode = "premask" if "premask" in short else ("fwd" if "_kernel<2, 0," in short else "bwd")
    launch.append(dict(kernel=name, mode=mode,
                       dur_us=val(r, "gpu__time_duration.sum"), dram_read_MB=val(r, "dram__bytes_read.sum"),
                       dram_write_MB=val(r, "dram__bytes_write.sum")))
# unit handling: ncu prints Mbyte / us for these sizes (checked in the header row)
units = rows[1]
assert units[idx["dram__bytes_read.sum"]] == "Mbyte" and units[idx["gpu__time_duration.sum"]] in ("us", "usecond"), units[idx["dram__bytes_read.sum"]]
fwd = [l for l in launch if l["mode"] == "fwd"]; bwd = [l for l in launch if l["mode"] == "bwd"]
avg = lambda ls: sum((l["dram_read_MB"] + l["dram_write_MB"]) for l in ls) / len(ls) * 1e6
traffic = {"gowalla": {"fwd": avg(fwd), "bwd": avg(bwd), "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full, cold-cache replay)",
                       "launches": launch}}
json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, os.path.join(P, "r1_ncu_full_summary.md")], stdout=subprocess.DEVNULL)
import shutil
shutil.copy(os.path.join(G, "launches_r1.csv"), os.path.join(P, "r1_launches.csv"))
for f in ("bench_full.json", "bench_ref.json"):
    shutil.copy(os.path.join(G, f), os.path.join(P, "r1_" + f))
print(json.dumps({k: traffic["gowalla"][k] for k in ("fwd", "bwd")}))
