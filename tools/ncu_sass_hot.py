"""Per-SASS-instruction executed counts and stall samples of one kernel from an `ncu --set full --import-source on`
report (source page, SASS view).  usage: python tools/ncu_sass_hot.py REPORT KERNEL_REGEX [min_share_pct]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout.splitlines()
# the report may hold several launches: take the first table
rows = list(csv.reader(out))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    body.append(r)
tot_inst = sum(float(r[col["Instructions Executed"]]) for r in body)
tot_samp = sum(float(r[col["# Samples"]]) for r in body)
print(f"{len(body)} SASS instructions, {tot_inst/1e6:.2f} M warp instructions, {tot_samp:.0f} samples")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(float(r[col[h]]) for r in body) for h in stalls}
print("stall samples:", {k[6:]: int(v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0})
cum = 0.0
for i, r in enumerate(body):
    inst = float(r[col["Instructions Executed"]])
    share = 100 * inst / tot_inst
    cum += share
    if share >= min_share:
        top = sorted(((float(r[col[h]]), h[6:]) for h in stalls), reverse=True)[:2]
        ts = " ".join(f"{n}:{int(v)}" for v, n in top if v > 0)
        print(f"{i:4d} {share:5.2f}% cum {cum:6.2f}% smp {int(float(r[col['# Samples']])):5d} thr {r[col['Avg. Predicated-On Threads Executed']]:>5s}  {r[col['Source']].strip()[:90]:90s} {ts}")
