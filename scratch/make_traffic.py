"""scratch: profiles/traffic.json entries from ncu captures.
    python scratch/make_traffic.py KEY REP.ncu-rep KERNEL_SUBSTRING [INDEX] [rays=N] [bound=hbm|issue] [note=...]
writes profiles/<rep basename>_raw.txt (the summary the numbers come from) and updates profiles/traffic.json[KEY]."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
key, rep, pat = sys.argv[1], sys.argv[2], sys.argv[3]
idx = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 0
kv = dict(a.split("=", 1) for a in sys.argv[4:] if "=" in a)
out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr = rows[0]
cands = [r for r in rows[2:] if pat in r[hdr.index("Kernel Name")]]
r = cands[idx]
g = lambda n: float(r[hdr.index(n)].replace(",", "")) if n in hdr and r[hdr.index(n)] not in ("", "n/a") else None
unit = lambda n: rows[1][hdr.index(n)]
def to_bytes(n):
    v = g(n); u = unit(n)
    return None if v is None else v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
dur = g("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit("gpu__time_duration.sum"), 1e-6)   # ms
base = os.path.basename(rep).replace(".ncu-rep", "").replace("_rawpage.csv", "")
summary = os.path.join(ROOT, "profiles", base + "_raw.txt")
with open(summary, "a") as f:
    f.write("---- %s  [%s, launch %d of the capture]\n" % (r[hdr.index("Kernel Name")], key, idx))
    for n in hdr:
        if any(n.startswith(p) for p in ("gpu__time_duration", "dram__bytes", "lts__t_sectors", "lts__t_sector_hit", "l1tex__t_sector_hit", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle",
                                           "smsp__thread_inst_executed_per", "sm__warps_active", "launch__", "sm__cycles", "smsp__average_warps_issue_stalled", "smsp__issue_active", "lts__throughput", "dram__throughput")):
            f.write("  %-96s %18s %s\n" % (n, r[hdr.index(n)], rows[1][hdr.index(n)]))
e = {"kernel": r[hdr.index("Kernel Name")][:100], "ms_in_capture": dur,
     "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
     "lts_bytes_per_launch": g("lts__t_sectors.sum") * 32.0, "lts_pct_of_peak": g("lts__t_sectors.sum.pct_of_peak_sustained_elapsed"),
     "warp_instructions_per_launch": g("smsp__inst_executed.sum"), "ipc": g("sm__inst_executed.avg.per_cycle_active"),
     "lanes_per_instruction": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
     "long_scoreboard_stall": g("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
     "l1_hit_pct": g("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": g("lts__t_sector_hit_rate.pct"),
     "source": "profiles/%s_raw.txt (ncu --set full --clock-control none; %s)" % (base, os.path.basename(rep).replace("_rawpage.csv", ".ncu-rep"))}
if "rays" in kv: e["rays_per_launch"] = int(float(kv["rays"]))
if "bound" in kv: e["bound"] = kv["bound"]
if "note" in kv: e["note"] = kv["note"]
p = os.path.join(ROOT, "profiles", "traffic.json")
d = json.load(open(p)) if os.path.exists(p) else {}
d[key] = e
json.dump(d, open(p, "w"), indent=1)
print(key, json.dumps(e, indent=1))
