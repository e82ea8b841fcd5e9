#!/usr/bin/env python3
"""Per-kernel totals and shares out of an ncu launch list (--metrics gpu__time_duration.sum --csv).
Usage: tools/launch_shares.py gpurun_out/x_launches.csv > profiles/x_launch_shares.txt"""
import collections, csv, sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i + 1
            break
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("av1b::<unnamed>::", "").replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        tot[name] += v
        cnt[name] += 1
    s = sum(tot.values())
    print("%-40s %6s %12s %10s %7s" % ("kernel", "n", "total us", "avg us", "share"))
    for k, v in tot.most_common():
        print("%-40s %6d %12.1f %10.1f %6.1f%%" % (k[-40:], cnt[k], v, v / cnt[k], 100 * v / s))
    print("%-40s %6d %12.1f" % ("total", sum(cnt.values()), s))


if __name__ == "__main__":
    main()
