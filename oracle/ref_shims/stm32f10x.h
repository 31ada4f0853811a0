/* ORACLE build shim (test infrastructure). The reference's VAD.C/MFCC.C/DTW.C only need the
 * fixed-width typedefs of the vendor header (Src/CM3_SYS stm32f10x.h:421-439: s8..u32 are the
 * <stdint.h> widths). Nothing else of the HAL is reproduced. */
#ifndef SR_SHIM_STM32F10X_H
#define SR_SHIM_STM32F10X_H
#include <stdint.h>
typedef int32_t  s32;
typedef int16_t  s16;
typedef int8_t   s8;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t  u8;
#endif
