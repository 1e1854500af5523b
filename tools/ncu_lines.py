"""Aggregate an ncu source page (CUDA view) per source line: samples and instructions executed.

    python tools/ncu_lines.py prof.ncu-rep [kernel-regex] [top N]
"""
import csv, io, subprocess, sys
path = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 2 and sys.argv[2]:
    cmd += ["-k", "regex:" + sys.argv[2]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = ""
hdr = None
acc = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 10 and r[0].isdigit():
        try:
            acc.append((int(r[hdr["# Samples"]]), int(r[hdr["Instructions Executed"]]), cur_file, int(r[0]), r[1].strip()[:110]))
        except ValueError:
            pass
ts = sum(a[0] for a in acc) or 1
ti = sum(a[1] for a in acc) or 1
print("total samples %d, total warp-instructions %d" % (ts, ti))
for s, i, f, ln, src in sorted(acc, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * s / ts, 100.0 * i / ti, f, ln, src))
