"""Per-kernel SASS evidence for the shipped library: how many tcgen05 MMA (UTCHMMA / UTCQMMA), TMEM load / store
(LDTM / STTM), TMA (UTMALDG / UTMASTG), ldmatrix (LDSM) and legacy tensor-core (HMMA) instructions each kernel of
libb200whisper.so contains.  Usage: python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "whisper-streaming-stt-server_b200", "libb200whisper.so")
OPS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "LDSM", "HMMA", "LDGSTS", "MUFU.EX2", "SYNCS", "UCGABAR", "ACQBULK")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = {}
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_n"] += 1
        for o in OPS:
            if op.startswith(o):
                counts[cur][o] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
arch = re.search(r"arch = (sm_\w+)", sass)
print(f"# cuobjdump -sass of {os.path.relpath(LIB, ROOT)} ({arch.group(1) if arch else '?'}): instruction counts per kernel")
print(f"# {'kernel':70s} {'instrs':>7s} " + " ".join(f"{o:>8s}" for o in OPS))
for (mangled, c), name in sorted(zip(counts.items(), dem), key=lambda t: t[1]):
    short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("bw::", "").replace("void ", "")
    print(f"{short[:72]:72s} {c['_n']:7d} " + " ".join(f"{c[o]:8d}" for o in OPS))
