"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) of tools/profile_stages.py:
per-kernel totals of the second encoder forward and of the last decoder step.  Times are cold-cache and
serialised (compare SHARES, not absolutes).  Usage: python profiles/analyze_launches.py <launches.csv>"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        val = val / 1e3 if r["Metric Unit"] == "ns" else val * 1e3 if r["Metric Unit"] == "ms" else val
        rows.append((int(r["ID"]), r["Kernel Name"], r["Grid Size"], val))
    return rows


def short(n):
    return re.sub(r"\(.*", "", n).replace("bw::", "").replace("<unnamed>::", "").replace("unnamed>::", "")[:64]


def table(title, rows):
    tot = sum(r[3] for r in rows)
    print(f"{title}: total {tot:.1f} us over {len(rows)} launches")
    a = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        k = (short(r[1]), r[2])
        a[k][0] += 1
        a[k][1] += r[3]
    for k, v in sorted(a.items(), key=lambda kv: -kv[1][1]):
        print(f"  {v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  x{v[0]:4d}  avg {v[1] / v[0]:8.1f} us  {k[0]}  grid {k[1]}")


rows = load(sys.argv[1])
first_dec = next(i for i, r in enumerate(rows) if "dec_embed" in r[1])
enc = rows[:first_dec]
table("ENCODER forward (2nd of 2)", enc[len(enc) // 2:])
decs = [i for i, r in enumerate(rows) if "dec_embed" in r[1]]
last = rows[decs[-1]:]
end = next(i for i, r in enumerate(last) if "beam_update" in r[1]) + 1
table("DECODER step (last)", last[:end])
