/* compat/ADC.H -- sampling constants of Src/BSP/ADC.H:7-11 (the capture itself is out of scope) */
#ifndef SR_COMPAT_ADC_H
#define SR_COMPAT_ADC_H
#include "stm32f10x.h"
#define fs         8000                     /* ADC.H:7  */
#define voice_len  2000                     /* ADC.H:8  */
#define VcBuf_Len  ((fs/1000)*voice_len)    /* ADC.H:9  */
#define atap_len_t 300                      /* ADC.H:10 */
#define atap_len   ((fs/1000)*atap_len_t)   /* ADC.H:11 */
#endif
