#!/usr/bin/env python
"""Per-source-line hot spots of one kernel from an .ncu-rep (source page, cuda+sass correlation).
usage: python tools/ncu_lines.py <rep> [top]   (needs ncu on PATH; the capture must have --import-source on)"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = []; cur = None; hdr = None; seen = set()
for r in csv.reader(out.splitlines()):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if len(r) == 2 and r[0] == "Kernel Name":
        if r[1] in seen: break      # first launch only
        seen.add(r[1]); continue
    if not hdr or len(r) < 10 or not r[0]: continue
    try: ln = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr[4:], r[4:]))
    def g(k):
        try: return float(d.get(k) or 0)
        except ValueError: return 0.0
    rows.append((cur, ln, r[1].strip()[:100], g("# Samples"), g("Instructions Executed"), g("Thread Instructions Executed")))
ts = sum(x[3] for x in rows) or 1; ti = sum(x[4] for x in rows) or 1
print("samples %d  warp-inst %d  avg lanes %.1f" % (ts, ti, sum(x[5] for x in rows) / ti))
for x in sorted(rows, key=lambda x: -x[3])[:top]:
    print("%-14s %4d smp=%5.1f%% inst=%5.1f%% lanes=%4.1f  %s" % (x[0], x[1], 100 * x[3] / ts, 100 * x[4] / ti, x[5] / x[4] if x[4] else 0, x[2]))
