"""Per-source-line instruction / stall-sample shares of one kernel in an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", f"regex:{kernel}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
ix, sm = h.index("Instructions Executed"), h.index("# Samples")
lines = []
for row in rows[hi + 1:]:
    if row and row[0] not in ("", "Line No", "File Path", "Function Name"):
        try:
            lines.append((int(row[0]), row[1], int(row[ix]), int(row[sm])))
        except ValueError:
            pass
tot = sum(l[2] for l in lines) or 1
ts = sum(l[3] for l in lines) or 1
print(f"{kernel}: {tot} warp instructions, {ts} samples")
for l in sorted(lines, key=lambda l: -l[2])[:top]:
    print(f"{l[0]:4d} inst {l[2] / tot * 100:5.1f}%  samples {l[3] / ts * 100:5.1f}%  {l[1][:110]}")
