"""Print the per-kernel times of the last repetition in an ncu `--metrics gpu__time_duration.sum --csv` log."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi, ui, gi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Grid Size")
out = [(r[ki][:90], float(r[vi].replace(",", "")), r[ui], r[gi]) for r in rows[hdr + 1:] if len(r) > vi]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = len(out) // reps
tot = 0.0
for o in out[-n:]:
    tot += o[1]
    print(f"{o[1]/1e3:10.1f} us  {o[0]}  grid {o[3]}")
print(f"{tot/1e3:10.1f} us  total of {n} launches")
