"""Per-kernel counts of the SASS mnemonics that show which hardware path a kernel uses (B200_PROFILING.md):
UTC*MMA = tcgen05.mma, UTMALDG/UTMASTG = TMA, LDTM/STTM = tcgen05.ld/st, HMMA = mma.sync, MUFU, DFMA ...

    python tools/sass_excerpt.py > profiles/r2_sass_excerpt.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spiral-tensorflow_b200", "csrc", "libspiral_b200.so")
PATS = collections.OrderedDict([("UTCHMMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"), ("UTMALDG (TMA load)", r"\bUTMALDG"),
                                ("UTMASTG (TMA store)", r"\bUTMASTG"), ("LDTM (tcgen05.ld)", r"\bLDTM"),
                                ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"), ("HMMA (mma.sync)", r"\bHMMA"),
                                ("LDGSTS (cp.async)", r"\bLDGSTS"), ("LDSM (ldmatrix)", r"\bLDSM"), ("MUFU", r"\bMUFU"), ("DFMA/DMUL/DADD", r"\bD(FMA|MUL|ADD)"),
                                ("instructions", r"^\s+/\*[0-9a-f]{4,5}\*/")])


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, rows = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name).replace("spiral::", "")
            rows[cur] = collections.Counter()
            continue
        if cur:
            for k, p in PATS.items():
                if re.search(p, line):
                    rows[cur][k] += 1
    print("# SASS excerpt of libspiral_b200.so (cuobjdump -sass, sm_100a) -- mnemonic counts per kernel\n")
    print("| kernel | " + " | ".join(PATS) + " |")
    print("|---|" + "---|" * len(PATS))
    tot = collections.Counter()
    for name, c in rows.items():
        print("| `%s` | " % name + " | ".join(str(c[k]) if c[k] else "" for k in PATS) + " |")
        tot.update(c)
    print("| **total** | " + " | ".join(str(tot[k]) for k in PATS) + " |")


if __name__ == "__main__":
    main()
