/* sr_compat.h -- one header carrying every name a host program written against the reference's
 * VAD.H / MFCC.H / DTW.H / Flash.H / ADC.H / stm32f10x.h expects, defined on top of ../speech_recog.h.
 * The same-named files next to this one only forward here, so a reference caller compiles unchanged with
 * -Iinclude/compat. Values cite where the reference defines them; nothing else of those headers is reproduced. */
#ifndef SR_COMPAT_H_
#define SR_COMPAT_H_
#include <stdint.h>
#include "../speech_recog.h"      /* atap_tag, valid_tag, v_ftr_tag, noise_atap(), VAD(), get_mfcc(), dtw() */

/* fixed-width aliases used throughout the reference (vendor stm32f10x.h:421-439) */
typedef int32_t s32;
typedef int16_t s16;
typedef int8_t s8;
typedef uint32_t u32;
typedef uint16_t u16;
typedef uint8_t u8;

/* sampling (Src/BSP/ADC.H:7-11) */
#define fs            SR_FS
#define voice_len     2000
#define VcBuf_Len     SR_VCBUF_LEN
#define atap_len_t    300
#define atap_len      SR_ATAP_LEN
/* framing (Src/Speech_Recog/VAD.H:4-8) */
#define max_vc_con    3
#define frame_time    20
#define frame_mov_t   10
#define frame_len     160
#define frame_mov     80
/* MFCC geometry (Src/Speech_Recog/MFCC.H:8-16) */
#define fft_point     1024
#define frq_max       512
#define tri_num       24
#define mfcc_num      12
#define vv_tim_max    1200
#define vv_frm_max    119
/* DTW sentinels (Src/Speech_Recog/DTW.H:4-5) */
#define dis_err       0xFFFFFFFF
#define dis_max       0xFFFFFFFF
/* template bank layout (Src/BSP/Flash.H:8-18); the flash driver itself is out of scope */
#define Flash_Fail     3
#define Flash_Success  0
#define save_mask      12345
#define size_per_ftr   4096
#define ftr_per_comm   4
#define size_per_comm  (ftr_per_comm * size_per_ftr)
#define comm_num       20
#define ftr_total_size (size_per_comm * comm_num)
#endif
