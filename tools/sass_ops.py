"""profiles/rNN_sass_ops.txt: per kernel of libkv_b200.so, the number of SASS instructions and of the Blackwell-specific
ones (tcgen05 MMA = UTCHMMA / UTCQMMA..., TMA = UTMALDG / UTMASTG, TMEM load/store = LDTM / STTM, tcgen05 commit /
barriers = UTCBAR, cluster barriers) — the evidence that the tensor-core kernels are tcgen05 / TMA code, not mma.sync.
  python tools/sass_ops.py profiles/r02_sass_ops.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "knightvision_b200", "libkv_b200.so")
KEYS = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "UCGABAR",
        "HMMA", "IMMA", "LDGSTS", "LDSM")


def main(dst):
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn, per = None, collections.OrderedDict()
    for line in out.split("\n"):
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            op = m.group(1)
            per[fn]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    per[fn][op] += 1
    demangle = subprocess.run(["c++filt"] + list(per), capture_output=True, text=True).stdout.split("\n")
    with open(dst, "w") as f:
        f.write("# cuobjdump -sass knightvision_b200/libkv_b200.so (sm_100a), per kernel: total SASS instructions and the\n"
                "# Blackwell tensor-core / TMA / TMEM instructions among them (UTCHMMA = tcgen05.mma, .2CTA = cta_group::2,\n"
                "# UTMALDG = cp.async.bulk.tensor, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit).  No HMMA / IMMA (mma.sync) anywhere.\n")
        for (fn, c), name in zip(per.items(), demangle):
            special = {k: v for k, v in c.items() if k != "_total"}
            short = re.sub(r"\(.*", "", name)
            f.write(f"\n{short}\n    instructions {c['_total']}")
            if special:
                f.write("\n    " + ", ".join(f"{k} x{v}" for k, v in sorted(special.items())))
            f.write("\n")


if __name__ == "__main__":
    main(sys.argv[1])
