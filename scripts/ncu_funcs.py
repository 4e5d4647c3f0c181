#!/usr/bin/env python3
"""Per-function and per-line view of an .ncu-rep (source-correlated): stall samples, executed warp instructions,
shared-memory wavefronts, and the executed opcode mix.  Usage: ncu_funcs.py report.ncu-rep [n_lines]"""
import bisect, collections, csv, io, os, re, subprocess, sys

rep = sys.argv[1]
nlines = int(sys.argv[2]) if len(sys.argv) > 2 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FUNC = re.compile(r"(?:LMH_HDN?|__global__|__device__(?: __forceinline__)?|template.*?LMH_HDN?)\s+.*?(\w+)\(")

def run(*a):
    return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout

sass = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "sass"))))
h = sass[1]; ix = {k: i for i, k in enumerate(h)}
ops = collections.Counter(); tot = 0
for r in sass[2:]:
    if len(r) < len(h): continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    if not m: continue
    op = m.group(2); b = op.split(".")[0]
    if b in ("LDS", "STS", "LDG", "STG", "LDL", "STL"): b = ".".join(op.split(".")[:3])[:10]
    n = int(r[ix["Instructions Executed"]]); ops[b] += n; tot += n
print("executed warp instructions:", tot)
print("  " + "  ".join(f"{k} {100 * v / tot:.1f}%" for k, v in ops.most_common(24)))

src = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "cuda,sass"))))
sections = [i for i, r in enumerate(src) if r and r[0] == "File Path"]
res = []
for si, start in enumerate(sections):
    path = src[start][1]
    end = sections[si + 1] if si + 1 < len(sections) else len(src)
    hh = src[start + 2]
    if "# Samples" not in hh: continue
    iS, iI, iW = hh.index("# Samples"), hh.index("Instructions Executed"), hh.index("L1 Wavefronts Shared")
    lp = path if os.path.exists(path) else os.path.join(ROOT, path.split("/root/repo/")[-1])
    text = open(lp).read().split("\n") if os.path.exists(lp) else []
    funcs = [(n, m.group(1)) for n, l in enumerate(text, 1) for m in [FUNC.match(l)] if m]
    for r in src[start + 3:end]:
        if r and r[0].isdigit():
            ln = int(r[0]); k = bisect.bisect_right([f[0] for f in funcs], ln) - 1
            try: res.append((os.path.basename(path), funcs[k][1] if k >= 0 else "?", ln, int(r[iS]), int(r[iI]), int(r[iW]), r[1]))
            except ValueError: pass
T = [sum(x[j] for x in res) or 1 for j in (3, 4, 5)]
agg = collections.defaultdict(lambda: [0, 0, 0])
for f, fn, ln, s, i, w_, _ in res:
    a = agg[(f, fn)]; a[0] += s; a[1] += i; a[2] += w_
print("\nfunction: stall samples / warp instructions / shared wavefronts (share of the kernel)")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:32]:
    print(f"  {k[0]:16s}{k[1]:28s} {100 * a[0] / T[0]:5.1f}% {100 * a[1] / T[1]:5.1f}% {100 * a[2] / T[2]:5.1f}%")
print("\nhottest lines")
for f, fn, ln, s, i, w_, txt in sorted(res, key=lambda x: -x[3])[:nlines]:
    print(f"  {f}:{ln:<5d} {100 * s / T[0]:4.1f}% {100 * i / T[1]:4.1f}% {100 * w_ / T[2]:4.1f}%  {txt.strip()[:110]}")
