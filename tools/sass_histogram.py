"""SASS instruction histogram per kernel of the built library -> profiles/r2_sass_histogram.txt.

    python tools/sass_histogram.py [lib.so] > profiles/r2_sass_histogram.txt
"""
import collections
import re
import subprocess
import sys

COLS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "FFMA2", "FMUL2",
        "FADD2", "FFMA", "MUFU", "HMMA", "DFMA", "SHFL", "LDS", "STS", "LDG", "STG", "ATOMS", "BAR"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else "wenet_b200/lib/libwenet_b200.so"
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.split("\n")
    print("# SASS instruction histogram per kernel of %s (cuobjdump -sass, sm_100a), round 2; tools/sass_histogram.py." % lib)
    print("# UTCHMMA = tcgen05.mma (a '.2CTA' suffix = cta_group::2), LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG/UTMAREDG = TMA "
          "tensor load/store/reduce, UBLKCP = cp.async.bulk,")
    print("# SYNCS = mbarrier ops, FFMA2/FMUL2/FADD2 = packed fp32x2.  HMMA (mma.sync) must be 0 everywhere.")
    print("kernel | total | " + " ".join(COLS) + " | UTCHMMA.2CTA")
    blocks = re.split(r"\n\s*Function : \S+\n", "\n" + sass)[1:]
    for name, body in zip(names, blocks):
        short = re.sub(r"\(.*$", "", name.replace("(anonymous namespace)::", "").replace("wb::", ""))
        ops = collections.Counter()
        total = pair = 0
        for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)((?:\.[A-Za-z0-9_]+)*)", body):
            total += 1
            ops[m.group(1)] += 1
            if m.group(1) == "UTCHMMA" and "2CTA" in m.group(2):
                pair += 1
        cells = " ".join("%s=%d" % (c, ops[c]) for c in COLS if ops[c])
        print("%s | %d | %s | %d" % (short, total, cells, pair))


if __name__ == "__main__":
    main()
