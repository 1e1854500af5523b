"""DRAM traffic and a few headline counters of the K1 kernels from ncu --set full captures -> profiles/r2_k1_traffic.json
(bench.py reads it for roofline.traffic) and a markdown summary on stdout.

    python tools/ncu_traffic.py --canvases 65536 --channels 3 --raster_math fast \
        --expand gpurun_out/r2_expand_b65536.ncu-rep --composite gpurun_out/r2_composite_b65536.ncu-rep
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    res = []
    for v in vals:
        d = {"kernel": v[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = (float(v[i].replace(",", "")), units[i])
                except ValueError:
                    d[k] = (v[i], units[i])
        res.append(d)
    return res


def to_bytes(val_unit):
    v, u = val_unit
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--canvases", type=int, required=True)
    ap.add_argument("--channels", type=int, default=3)
    ap.add_argument("--raster_math", default="fast")
    ap.add_argument("--expand", required=True)
    ap.add_argument("--composite", required=True)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_k1_traffic.json"))
    a = ap.parse_args()
    rows = {}
    for name, path in (("expand", a.expand), ("composite", a.composite)):
        caps = raw(path)
        # average over the captured launches
        agg = {}
        for k in KEYS:
            xs = [c[k] for c in caps if k in c and isinstance(c[k][0], float)]
            if xs:
                agg[k] = (sum(x[0] for x in xs) / len(xs), xs[0][1])
        rows[name] = agg
        print("### %s_kernel (%d launches captured, B=%d, %s)" % (name, len(caps), a.canvases, a.raster_math))
        for k, (v, u) in agg.items():
            print("- %s: %.4g %s" % (k, v, u))
        print()
    entry = dict(canvases=a.canvases, channels=a.channels, raster_math=a.raster_math,
                 expand_bytes=to_bytes(rows["expand"]["dram__bytes_read.sum"]) + to_bytes(rows["expand"]["dram__bytes_write.sum"]),
                 composite_bytes=to_bytes(rows["composite"]["dram__bytes_read.sum"]) + to_bytes(rows["composite"]["dram__bytes_write.sum"]),
                 source="ncu --set full --clock-control none, %s + %s (dram__bytes_read.sum + dram__bytes_write.sum per launch)"
                        % (os.path.basename(a.expand), os.path.basename(a.composite)))
    old = []
    if os.path.exists(a.out):
        with open(a.out) as fp:
            old = [r for r in json.load(fp) if not (r["canvases"] == a.canvases and r["channels"] == a.channels and r["raster_math"] == a.raster_math)]
    with open(a.out, "w") as fp:
        json.dump(old + [entry], fp, indent=1)
    print("wrote", a.out, entry)


if __name__ == "__main__":
    main()
