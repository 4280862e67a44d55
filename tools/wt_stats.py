"""Median per-phase SM clocks of cmux_wide from a -DSPF_WIDE_TRACE build's device printf lines (stdin)."""
import sys, statistics, collections
cols = collections.defaultdict(list)
for line in sys.stdin:
    if not line.startswith("WT "): continue
    t = line.split()[1:]
    for k, v in zip(t[0::2], t[1::2]): cols[k].append(int(v))
for k, v in cols.items(): print(f"{k:6s} median {statistics.median(v):8.0f} clk  p10 {sorted(v)[len(v)//10]:8d}  p90 {sorted(v)[9*len(v)//10]:8d}  n {len(v)}")
