/* ORACLE build shim. VAD.C:4 / MFCC.C:5 / DTW.C:5 include "ADC.h" (lower case) while the file in
 * the reference tree is Src/BSP/ADC.H; on a case-sensitive host forward to it by absolute path so
 * that the reference's own fs / VcBuf_Len / atap_len (ADC.H:7-11) are used unmodified.
 * (Src/BSP is NOT put on the include path: it carries a private stdint.h.) */
#include SR_REF_ADC_H
