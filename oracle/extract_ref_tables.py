#!/usr/bin/env python3
"""ORACLE build helper: pull the 2040 twiddle halfwords out of the reference's
Src/BSP/cr4_fft_1024_stm32.s:285-629 (DCW lines of TableFFT_V7) into oracle/_ref/twiddle_ref.h.
Output goes only into oracle/_ref/ (git-ignored); no reference source is copied into the repo."""
import re
import sys

src_path, dst_path = sys.argv[1], sys.argv[2]
text = open(src_path, "rb").read().decode("gb18030").replace("\r", "")
body = text[text.index("TableFFT_V7\n"):]
vals = []
for line in body.split("\n"):
    if "DCW" in line:
        vals += [int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{4})", line)]
assert len(vals) == 2040, len(vals)
vals = [v - 65536 if v >= 32768 else v for v in vals]
with open(dst_path, "w") as f:
    f.write("/* extracted from %s -- build artefact, not committed */\n" % src_path)
    f.write("#include <stdint.h>\nstatic const int16_t sr_ref_twiddle[2040] = {\n")
    for i in range(0, 2040, 12):
        f.write("  " + ",".join(str(v) for v in vals[i:i + 12]) + ",\n")
    f.write("};\n")
