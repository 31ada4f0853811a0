#!/bin/bash
# ORACLE hygiene check (needs /root/reference and a built oracle/_ref): rebuild the reference's own VAD.C/MFCC.C/DTW.C with
# -fsanitize=address,undefined and run the golden captures + a seeded synthetic batch through it, proving that the inputs
# used for pinning never reach the reference's latent hazards (vc_dat[-1] read MFCC.C:119, log(0) MFCC.C:168, signed
# overflow). Output only under /tmp. Usage: bash oracle/asan_check.sh
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${REF:-/root/reference}
OUT=/tmp/sr_asan; mkdir -p $OUT
gcc -std=gnu99 -O1 -g -fno-omit-frame-pointer -fsanitize=address,undefined -ffp-contract=off -fPIC -w -shared -o $OUT/libref.so \
  -I$HERE/ref_shims -I$REF/Src/Speech_Recog -I$HERE/_ref \
  -DSR_REF_ADC_H="\"$REF/Src/BSP/ADC.H\"" -DSR_REF_FLASH_H="\"$REF/Src/BSP/Flash.H\"" \
  -DSR_TWIDDLE_HEADER='"twiddle_ref.h"' -DSR_TWIDDLE_NAME=sr_ref_twiddle \
  -x c $REF/Src/Speech_Recog/VAD.C -x c $REF/Src/Speech_Recog/MFCC.C -x c $REF/Src/Speech_Recog/DTW.C \
  -x c $HERE/ref_driver.c -x c $HERE/cr4_fft_restated.c -lm
cat > $OUT/run.py <<PY
import sys, numpy as np
sys.path.insert(0, "$HERE/../tests"); sys.path.insert(0, "$HERE/../stm32-speech-recognition_b200/python")
import oracle_bind as ob, sr_b200
ob.REF_SO = "$OUT/libref.so"
r = ob.RefOracle()
caps = np.load("$HERE/../tests/golden/captures.npz")
for name in caps.files:
    pcm = caps[name]; a = r.noise_atap(pcm, 2400); seg = r.vad(pcm, len(pcm), a)
    for k in range(3):
        if seg[2 * k + 1] != ob.NULL:
            r.mfcc_batch(pcm.reshape(1, -1), seg[2 * k:2 * k + 2].reshape(1, 2), a)
pcm = sr_b200.synth_pcm_host(256, 8000, 0x5EED0000); tpl = sr_b200.synth_pcm_host(8, 8000, 0x7E3A0000)
bank = sr_b200.make_bank(r.recognise_batch(tpl, 2400, None, 0, 4096)["ftr"])
o = r.recognise_batch(pcm, 2400, bank, 8, 4096)
pcm5 = sr_b200.synth_pcm_host(8, 40000, 0x5EED5000, 3); r.recognise_batch(pcm5, 2400, bank, 8, 4096)
print("sanitizer run finished cleanly; status histogram", np.bincount(o["status"], minlength=3).tolist())
PY
LD_PRELOAD=$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
  UBSAN_OPTIONS=print_stacktrace=1 python $OUT/run.py
