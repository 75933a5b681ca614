"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into per-kernel shares.

  python tools/launch_shares.py gpurun_out/launches.csv > profiles/rNN_launch_shares.csv
Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes."""
import collections
import csv
import re
import sys


def short(name: str) -> str:
  name = re.sub(r"\(.*", "", name).replace("void ", "")
  if name.startswith("at::") or "at::native" in name:
    return "torch glue (elementwise / rand / copy kernels)"
  return name


def main(path: str):
  rows = [r for r in csv.reader(open(path)) if len(r) > 5]
  hdr = rows[0]
  ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
  tot = collections.defaultdict(float)
  cnt = collections.Counter()
  for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
    k = short(r[ki])
    tot[k] += us
    cnt[k] += 1
  total = sum(tot.values())
  print(f"# {path}: {sum(cnt.values())} launches, {total / 1e3:.2f} ms (cold-cache, serialised: compare shares)")
  print("kernel,launches,total_us,share_pct,avg_us")
  for k in sorted(tot, key=tot.get, reverse=True):
    print(f"{k},{cnt[k]},{tot[k]:.1f},{100 * tot[k] / total:.1f},{tot[k] / cnt[k]:.1f}")


if __name__ == "__main__":
  main(sys.argv[1])
