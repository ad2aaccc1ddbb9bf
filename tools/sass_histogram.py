"""SASS opcode histogram of the shipped libcfr_b200.so, per kernel (cuobjdump -sass): the evidence that the conv path
is tcgen05 / TMEM / TMA code (UTCHMMA = tcgen05.mma kind::f16, UTMALDG = cp.async.bulk.tensor, LDTM = tcgen05.ld,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, USETMAXREG = setmaxnreg) and carries no legacy HMMA.

    python tools/sass_histogram.py > profiles/sass_r02_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "certifyingfacerecognition_b200", "libcfr_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "USETMAXREG",
       "HMMA", "IMMA", "LDGSTS", "ATOMS", "ATOMG", "RED", "STG", "LDG", "STS", "LDS", "SHFL", "BAR", "FFMA", "HFMA2", "FADD2",
       "FFMA2", "FMUL2", "F2FP", "MUFU")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = per.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    stamp = ""
    sp = LIB + ".stamp"
    if os.path.isfile(sp):
        stamp = open(sp).read()[:16]
    print(f"# SASS opcode histogram of {os.path.relpath(LIB, ROOT)} (source fingerprint {stamp}), {len(per)} kernels, "
          f"{sum(total.values())} instructions")
    print("# whole library: " + "  ".join(f"{k}={total[k]}" for k in KEY if total[k]))
    print(f"# legacy tensor-core opcodes: HMMA={total['HMMA']} IMMA={total['IMMA']} (must be 0)")
    print("kernel\tinstructions\t" + "\t".join(KEY[:12]))
    for name, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values())):
        print(f"{name}\t{sum(c.values())}\t" + "\t".join(str(c[k]) for k in KEY[:12]))


if __name__ == "__main__":
    main()
