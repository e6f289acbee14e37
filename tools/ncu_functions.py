#!/usr/bin/env python3
"""Per-function share of pc-sampling stalls and executed instructions for one kernel of an .ncu-rep."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
seen, segs = set(), []
cur = {"n": 0, "samp": 0, "exec": 0, "imadw": 0, "ldst": 0, "first": None}
tot = 0
for r in rows[2:]:
    try:
        s, n = int(r[idx["# Samples"]]), int(r[idx["Instructions Executed"]])
    except (ValueError, IndexError):
        continue
    if r[idx["Address"]] in seen:
        continue
    seen.add(r[idx["Address"]])
    src = r[idx["Source"]].strip().split()
    op = src[1] if src[0].startswith("@") else src[0]
    if cur["first"] is None:
        cur["first"] = r[idx["Address"]]
    cur["n"] += 1; cur["samp"] += s; cur["exec"] += n; tot += s
    cur["imadw"] += op.startswith("IMAD.WIDE")
    cur["ldst"] += op.startswith(("LD", "ST"))
    if op.startswith(("RET", "EXIT")):
        segs.append(cur)
        cur = {"n": 0, "samp": 0, "exec": 0, "imadw": 0, "ldst": 0, "first": None}
texec = sum(s["exec"] for s in segs)
print("functions %d, samples %d, warp instructions %.3g" % (len(segs), tot, texec))
for s in sorted(segs, key=lambda x: -x["samp"])[:16]:
    print("  size %4d (imad.wide %3d, ld/st %3d)  samples %5.1f%%  instructions %5.1f%%  samples/instr %.2f" % (
        s["n"], s["imadw"], s["ldst"], 100 * s["samp"] / tot, 100 * s["exec"] / texec, (s["samp"] / tot) / max(s["exec"] / texec, 1e-9)))
