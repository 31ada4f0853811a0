/* compat/stm32f10x.h -- lets a host program written against the reference's headers compile unchanged:
 * only the fixed-width typedefs the hot path uses (vendor header stm32f10x.h:421-439). */
#ifndef SR_COMPAT_STM32F10X_H
#define SR_COMPAT_STM32F10X_H
#include <stdint.h>
typedef int32_t s32;
typedef int16_t s16;
typedef int8_t s8;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;
#endif
