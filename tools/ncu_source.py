#!/usr/bin/env python
"""Per-region hot spots from `ncu -i X.ncu-rep --page source --csv`: splits the SASS into
basic-block-like runs at backward branches / labels and prints, for the heaviest runs,
instructions executed, average active threads and stall samples."""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True,
                     text=True).stdout.splitlines()
rows = list(csv.reader(out))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot_inst = sum(float(r[ix["Instructions Executed"]] or 0) for r in data)
tot_samp = sum(float(r[ix["# Samples"]] or 0) for r in data)
print(f"instructions executed {tot_inst:.3e}   samples {tot_samp:.0f}")
# group in windows of consecutive instructions with similar execution counts
groups = []
cur = None
for n, r in enumerate(data):
    ie = float(r[ix["Instructions Executed"]] or 0)
    if cur is None or abs(ie - cur["ie0"]) > 0.02 * max(ie, cur["ie0"], 1):
        cur = dict(start=n, ie0=ie, inst=0.0, thr=0.0, samp=0.0, n=0, ops={}, stalls={})
        groups.append(cur)
    cur["inst"] += ie
    cur["thr"] += float(r[ix["Thread Instructions Executed"]] or 0)
    cur["samp"] += float(r[ix["# Samples"]] or 0)
    cur["n"] += 1
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"):
        op = r[ix["Source"]].split()[1]
    cur["ops"][op.split(".")[0]] = cur["ops"].get(op.split(".")[0], 0) + 1
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            v = float(r[ix[h]] or 0)
            if v:
                cur["stalls"][h] = cur["stalls"].get(h, 0) + v
top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
for g in sorted(groups, key=lambda g: -g["samp"])[:top]:
    ops = " ".join(f"{k}:{v}" for k, v in sorted(g["ops"].items(), key=lambda x: -x[1])[:7])
    st = " ".join(f"{k[6:]}:{v:.0f}" for k, v in sorted(g["stalls"].items(), key=lambda x: -x[1])[:4])
    print(f"@{g['start']:5d} n={g['n']:3d} exec/inst={g['ie0']:.3e} inst%={100 * g['inst'] / tot_inst:5.1f} "
          f"samp%={100 * g['samp'] / tot_samp:5.1f} act={g['thr'] / max(g['inst'], 1):5.1f} | {ops} | {st}")
