#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of libspeech_b200.so -> profiles/rN_sass_opcodes.md (evidence that the sm_100a
features the design names are really in the binary: UBLKCP bulk copies + SYNCS mbarriers, IDP dot products, packed
VIMNMX, no tensor-core / library code).  Usage: python tools/sass_histogram.py profiles/r2_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
LIB = os.path.join(ROOT, "stm32-speech-recognition_b200", "lib", "libspeech_b200.so")
WATCH = ["UBLKCP", "SYNCS", "IDP", "VIMNMX", "IMAD.HI", "IMAD.WIDE", "LEA.HI", "SHF", "PRMT", "MUFU", "REDS", "ATOMS", "LDS", "STS",
         "LDG", "STG", "SHFL", "HMMA", "UTMALDG", "UTCHMMA", "BAR"]


def main():
    out = sys.argv[1]
    sass = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    fn, hist = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0]
            hist[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            hist[fn][m.group(1)] += 1
    with open(out, "w") as f:
        f.write("# SASS opcode histogram of libspeech_b200.so (cuobjdump -sass, sm_100a), static instruction counts\n\n")
        f.write("| kernel | instr | " + " | ".join(WATCH) + " |\n|---|---|" + "---|" * len(WATCH) + "\n")
        for k, c in hist.items():
            tot = sum(c.values())
            cells = []
            for w in WATCH:
                n = sum(v for o, v in c.items() if o == w or o.startswith(w + ".") or (w in ("IMAD.HI", "IMAD.WIDE", "LEA.HI") and o.startswith(w)))
                cells.append(str(n) if n else "")
            f.write("| `%s` | %d | %s |\n" % (k.replace("srk::", ""), tot, " | ".join(cells)))
        f.write("\nTop opcodes per hot kernel:\n\n")
        for k, c in hist.items():
            if any(s in k for s in ("mfcc_kernel_s16", "dtw_kernel", "dtw_dyn_kernel", "vad_kernel", "stream_step_kernel")):
                f.write("* `%s`: " % k.replace("srk::", "") + ", ".join("%s %d" % (o, n) for o, n in c.most_common(14)) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
