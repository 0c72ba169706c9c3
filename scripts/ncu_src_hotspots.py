#!/usr/bin/env python
"""Per-source-line hotspots of one kernel from an .ncu-rep captured with --import-source on.
usage: ncu_src_hotspots.py prof.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = ""; hdr = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 4 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < 10 or r[0] == "": continue
    try:
        s = int(r[4]); inst = int(r[7]); thr = int(r[8])
    except ValueError:
        continue
    out.append((s, inst, thr, cur_file, r[0], r[1].strip()[:130]))
tot = sum(o[0] for o in out) or 1; toti = sum(o[1] for o in out) or 1
out.sort(reverse=True)
print("total stall samples %d, warp instructions %d" % (tot, toti))
print("samples%  inst%  lanes  file:line  source")
for s, inst, thr, f, ln, src in out[:top]:
    print("%5.1f  %5.1f  %4.1f  %s:%s  %s" % (100.0 * s / tot, 100.0 * inst / toti, thr / inst if inst else 0, f, ln, src))
