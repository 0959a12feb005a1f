"""Per-kernel SASS mnemonic counts of libbaler_b200.so (tensor-core, TMEM, bulk-copy, FMA instructions):
    python tools/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baler_b200", "libbaler_b200.so")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "LDSM", "LDGSTS", "FFMA", "DFMA", "SHFL",
         "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR", "SYNCS"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kern, counts = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        dem = dem.replace("(anonymous namespace)::", "")
        m2 = re.search(r"([A-Za-z_]\w*(?:<[^()]*>)?)\(", dem)
        kern = m2.group(1) if m2 else dem
        counts.setdefault(kern, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        counts[kern]["total"] += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w):
                counts[kern][w] += 1
                break
print("SASS of %s (sm_100a), instruction counts per kernel" % os.path.basename(LIB))
for k, c in sorted(counts.items(), key=lambda kv: -kv[1]["total"]):
    print("%-90s total %6d | %s" % (k[-90:], c["total"], " ".join("%s %d" % (w, c[w]) for w in WATCH if c[w])))
