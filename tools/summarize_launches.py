#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel.

    python tools/summarize_launches.py gpurun_out/launches.csv [-o profiles/r1_launches.md] [--title "..."]
"""
import argparse
import csv
import io
import re
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("s2m2::", "")
    name = re.sub(r"void ", "", name)
    return name.strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("-o", "--out", default=None)
    ap.add_argument("--title", default="ncu launch list")
    a = ap.parse_args()
    text = open(a.csv, errors="replace").read()
    start = text.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = defaultdict(lambda: [0, 0.0])
    unit = None
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1e-3)
        k = short(r["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += v * scale
    total = sum(v[1] for v in agg.values())
    lines = [f"# {a.title}", "", f"source: `{a.csv}` — {sum(v[0] for v in agg.values())} launches, {total / 1e3:.2f} ms summed "
             "(per-launch times under ncu are serialised / cold-cache: compare shares, not absolutes)", "",
             "| kernel | launches | total ms | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {t / 1e3:.3f} | {100 * t / total:.1f}% | {t / n:.1f} |")
    out = "\n".join(lines) + "\n"
    if a.out:
        open(a.out, "w").write(out)
    print(out)


if __name__ == "__main__":
    main()
