"""Opcode census of the in-tree libsgb200.so (the library is git-ignored, so the evidence that the contraction kernels are
tcgen05 / TMEM / TMA code is committed as text):  python profiles/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pytorch-studiogan_b200", "sgb200", "lib", "libsgb200.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "DFMA", "HMMA", "ATOMS", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            per[cur]["_total"] += 1
            if op in OPS:
                per[cur][op] += 1
    print("cuobjdump -sass pytorch-studiogan_b200/sgb200/lib/libsgb200.so  (sm_100a): instructions per kernel")
    print("%-44s %7s " % ("kernel", "total") + " ".join("%8s" % o for o in OPS))
    for k, c in per.items():
        print("%-44s %7d " % (k[-44:], c["_total"]) + " ".join("%8d" % c[o] for o in OPS))
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("%-44s %7d " % ("ALL", tot["_total"]) + " ".join("%8d" % tot[o] for o in OPS))


if __name__ == "__main__":
    sys.exit(main())
