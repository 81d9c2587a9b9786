"""Summarise an ncu launch list (gpu__time_duration.sum CSV) by kernel and grid."""
import collections
import csv
import re
import sys


def summarize(path, top=24):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= vi:
            continue
        key = (re.sub(r'\(.*', '', r[ki]).replace('void ', ''), r[gi])
        agg[key][0] += 1
        agg[key][1] += float(r[vi].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    out = [f"{n} launches, total {tot / 1e3:.0f} us (cold-cache, serialised: compare shares)", "",
           "| share | launches | avg us | kernel | grid |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f"| {v[1] / tot * 100:.1f}% | {v[0]} | {v[1] / v[0] / 1000:.1f} | `{k[0][:48]}` | {k[1]} |")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
