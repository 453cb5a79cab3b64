#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel launch counts and time shares."""
import collections, csv, re, sys

def main(path, out):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    ui = hdr.index("Metric Unit")
    tot = collections.Counter(); cnt = collections.Counter()
    for r in rd:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "s": 1e9, "second": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r[ki])[:90]
        tot[name] += ns; cnt[name] += 1
    total = sum(tot.values())
    lines = [f"# per-kernel shares of {path} (serialised, cold-cache ncu timings: compare SHARES)", f"total {total/1e6:.1f} ms over {sum(cnt.values())} launches", ""]
    for name, ns in tot.most_common(25):
        lines.append(f"{100*ns/total:6.2f}%  {ns/1e6:10.2f} ms  {cnt[name]:6d} launches  {name}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
