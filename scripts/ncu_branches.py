#!/usr/bin/env python3
"""Executed control-flow instructions (BRA/BSSY/BSYNC/BRX/CALL) and branch-resolving stall samples per source line
of an .ncu-rep.  Usage: ncu_branches.py report.ncu-rep [n]"""
import bisect, collections, csv, io, os, re, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
src = list(csv.reader(io.StringIO(out)))
sections = [i for i, r in enumerate(src) if r and r[0] == "File Path"]
ll = collections.Counter(); st = collections.Counter(); txt = {}; tot_i = 0
for si, start in enumerate(sections):
    path = src[start][1]; end = sections[si + 1] if si + 1 < len(sections) else len(src)
    hh = src[start + 2]
    if "# Samples" not in hh: continue
    iI = hh.index("Instructions Executed"); iB = hh.index("stall_branch_resolving")
    ln = 0
    for r in src[start + 3:end]:
        if not r: continue
        if r[0].isdigit():
            ln = int(r[0]); txt[(os.path.basename(path), ln)] = r[1].strip()
            try: st[(os.path.basename(path), ln)] += int(r[iB]); tot_i += int(r[iI])
            except ValueError: pass
        elif r[0] == "" and len(r) > iI and r[3] not in ("...", "-"):
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
            if m and m.group(2).split(".")[0] in ("BRA", "BSYNC", "BSSY", "BRX", "CALL", "RET"):
                try: ll[(os.path.basename(path), ln)] += int(r[iI])
                except ValueError: pass
tot = sum(ll.values()) or 1; ts = sum(st.values()) or 1
print(f"control instructions (non-elided rows): {tot} = {100 * tot / tot_i:.1f}% of executed; branch-resolving samples: {ts}")
for k, n in ll.most_common(N):
    print(f"  {k[0]}:{k[1]:<5d} {100 * n / tot:5.1f}% ctrl  {100 * st[k] / ts:5.1f}% br-stall  {txt.get(k, '')[:100]}")
