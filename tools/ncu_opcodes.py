#!/usr/bin/env python
"""Executed-instruction histogram by SASS opcode (first launch of an .ncu-rep captured with --import-source on).
usage: python tools/ncu_opcodes.py <rep> [top]"""
import csv, re, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
hdr = None; n = 0; h = collections.Counter(); lanes = collections.Counter(); smp = collections.Counter()
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        n += 1
        if n > 1: break
        continue
    if r and r[0] == "Address": hdr = r; continue
    if not hdr or len(r) < 8: continue
    d = dict(zip(hdr, r))
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", d["Source"])
    if not m: continue
    op = m.group(2)
    h[op] += float(d["Instructions Executed"] or 0); lanes[op] += float(d["Thread Instructions Executed"] or 0); smp[op] += float(d["# Samples"] or 0)
tot = sum(h.values()); ts = sum(smp.values()) or 1
print("warp-inst %d" % tot)
for op, c in h.most_common(top):
    print("%-10s inst=%5.1f%%  smp=%5.1f%%  lanes=%4.1f" % (op, 100 * c / tot, 100 * smp[op] / ts, lanes[op] / c if c else 0))
