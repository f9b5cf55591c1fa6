"""Per-kernel histogram of the Blackwell-specific SASS mnemonics in libobman_b200.so (cuobjdump -sass): UTC*MMA
(tcgen05.mma), UTMALDG / UTMASTG (TMA), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit), SYNCS (mbarrier), plus
HMMA as the tell-tale of a legacy tensor path.  `python scripts/sass_summary.py > profiles/sass_summary_<round>.txt`"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "obman_train_b200", "libobman_b200.so")
PATTERNS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "UTCBAR", "UTCATOM",
            "SYNCS", "HMMA", "LDGSTS", "ATOMG", "RED", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            kernels[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            kernels[name]["total"] += 1
            for p in PATTERNS:
                if op.startswith(p):
                    kernels[name][p] += 1
                    break
    print("# cuobjdump -sass obman_train_b200/libobman_b200.so (sm_100a): instruction counts per kernel")
    print("# UTC*MMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, SYNCS = mbarrier")
    cols = [p for p in PATTERNS if any(k[p] for k in kernels.values())]
    print("%-78s %7s " % ("kernel", "instr") + " ".join("%7s" % c for c in cols))
    tot = collections.Counter()
    for name, c in kernels.items():
        print("%-78s %7d " % (name[:78], c["total"]) + " ".join("%7d" % c[p] for p in cols))
        tot.update(c)
    print("%-78s %7d " % ("ALL %d kernels" % len(kernels), tot["total"]) + " ".join("%7d" % tot[p] for p in cols))


if __name__ == "__main__":
    sys.exit(main())
