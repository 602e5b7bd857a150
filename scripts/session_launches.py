"""One session's launches out of an ncu `--metrics gpu__time_duration.sum --csv` log of bench.py: the rows between two
consecutive starts of a plan (k_plan_edges), i.e. the plan of the next session + build + stats + aggregate + head + release
of this one, with the kernels' shares.  Of the windows that contain the head kernel (the end-to-end leg) the one with the
smallest total is printed: the first sessions on a stream also carry the one-time zero fills of its workspaces.  Times under
ncu are cold-cache and serialised: shares, not absolutes."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size")
out = [(r[ki], float(r[vi].replace(",", "")), r[gi]) for r in rows[hdr + 1:] if len(r) > vi]
starts = [i for i, o in enumerate(out) if "k_plan_edges" in o[0]]
wins = [(starts[i], starts[i + 1]) for i in range(len(starts) - 1)]
wins = [w for w in wins if any("k_cn_head" in o[0] for o in out[w[0]:w[1]])] or wins
a, b = min(wins, key=lambda w: sum(o[1] for o in out[w[0]:w[1]]))
tot = sum(o[1] for o in out[a:b])
for o in out[a:b]:
    print(f"{o[1] / 1e3:10.1f} us  {o[0][:110]}  grid {o[2]}")
print(f"{tot / 1e3:10.1f} us  total of {b - a} launches")
share = defaultdict(float)
for o in out[a:b]:
    share[o[0].split("(")[0].replace("void ", "").split("<")[0]] += o[1]
print("# shares by kernel")
for k, v in sorted(share.items(), key=lambda kv: -kv[1])[:12]:
    print(f"{v / tot * 100:6.1f} %  {v / 1e3:9.1f} us  {k}")
