#!/usr/bin/env python
"""Per-kernel SASS instruction summary of libpvs_b200.so (no GPU needed): counts of the mnemonics that prove the
tcgen05 / TMEM / TMA path (UTCHMMA / UTCQMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st,
UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy, SYNCS = mbarrier) next to HMMA / IMMA (legacy mma.sync,
expected to be zero) and the spill traffic (LDL / STL).

    python tools/sass_summary.py > profiles/sass_summary_r02.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "python-visual-similarity_b200", "pyvisim_b200", "lib", "libpvs_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "IMMA", "MUFU", "LDL", "STL"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
rows, cur, k = [], None, -1
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        k += 1
        cur = collections.Counter()
        rows.append((names[k] if k < len(names) else m.group(1), cur))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for w in WATCH:
            if op.startswith(w):
                cur[w] += 1
print(f"# {os.path.relpath(LIB, ROOT)}: {len(rows)} kernels, SASS for sm_100a (cuobjdump -sass)")
print("# " + " ".join(f"{w:>7}" for w in ["total"] + WATCH) + "  kernel")
tot = collections.Counter()
for name, c in sorted(rows, key=lambda r: -r[1]["UTCHMMA"] - r[1]["UTCQMMA"]):
    tot.update(c)
    short = re.sub(r"\(.*", "", name)[:110]
    print("  " + " ".join(f"{c[w]:7d}" for w in ["total"] + WATCH) + "  " + short)
print("# " + " ".join(f"{tot[w]:7d}" for w in ["total"] + WATCH) + "  ALL")
