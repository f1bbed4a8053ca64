#!/usr/bin/env python
"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel markdown table
(shares of the step; absolute times are cold-cache and serialised, see B200_PROFILING.md)."""
import collections
import csv
import sys


def main(path, out=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        u = row.get("Metric Unit", "us")
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        name = row["Kernel Name"].split("(")[0]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    rows = ["| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        rows.append("| `%s` | %d | %.1f | %.2f | %.1f%% |" % (k, cnt[k], v, v / cnt[k], 100 * v / T))
    rows.append("| **total** | %d | %.1f | | |" % (sum(cnt.values()), T))
    text = "\n".join(rows) + "\n"
    if out:
        open(out, "w").write(text)
    else:
        sys.stdout.write(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
