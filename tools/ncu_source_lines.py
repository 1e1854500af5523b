"""Per-source-line / per-phase instruction counts of one kernel from an ncu report captured with --import-source on.

    ncu -i REPORT.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all libspiral_b200.so ; nvdisasm -c -g raster.sm_100a.cubin > raster.disasm
    python tools/ncu_source_lines.py sass.csv raster.disasm '.text._ZN6spiral16composite_kernelILi3ELb0E' raster.cu

ncu's CSV source page carries SASS rows only; nvdisasm -g carries the `//## File ..., line N` markers of -lineinfo.  The two
list the kernel's instructions in the same order, so they are joined by position (opcodes are cross-checked)."""
import collections
import csv
import re
import sys


def main(sass_csv, disasm, symbol_prefix, src_name):
    lines = open(disasm).read().split("\n")
    start = [i for i, l in enumerate(lines) if l.startswith(symbol_prefix)][0]
    end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith("//--------------------- "))
    cur, ins = None, []
    for l in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((cur, m.group(2)))
    rows = list(csv.reader(open(sass_csv)))
    hdr, data = rows[1], rows[2:]
    ie, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    n = len(ins)
    op = lambda t: [w for w in t.split() if not w.startswith("@")][0].split(".")[0]
    bad = sum(1 for i in range(n) if op(ins[i][1]) != op(data[i][1]))
    agg = collections.defaultdict(lambda: [0, 0])
    for i in range(n):
        agg[ins[i][0] or ("?", 0)][0] += int(data[i][ie])
        agg[ins[i][0] or ("?", 0)][1] += int(data[i][isamp])
    tot, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
    print("%d SASS instructions, %d opcode mismatches, %d warp-instructions executed, %d stall samples" % (n, bad, tot, ts))
    for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:40]:
        print("%6.2f %% instr %6.2f %% samples  %s:%d" % (100.0 * v[0] / tot, 100.0 * v[1] / ts, f, ln))
    return agg, tot, ts


if __name__ == "__main__":
    main(*sys.argv[1:5])
