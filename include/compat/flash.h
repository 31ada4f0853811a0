/* compat/Flash.H -- template-bank layout constants of Src/BSP/Flash.H:11-20 (the flash driver is out of scope;
 * the bank lives in host/device memory with the same slot layout) */
#ifndef SR_COMPAT_FLASH_H
#define SR_COMPAT_FLASH_H
#include "MFCC.H"
#define save_mask      12345                /* Flash.H:11 */
#define size_per_ftr   (4*1024)             /* Flash.H:13 */
#define ftr_per_comm   4                    /* Flash.H:15 */
#define size_per_comm  (ftr_per_comm*size_per_ftr)
#define comm_num       20                   /* Flash.H:17 */
#define ftr_total_size (size_per_comm*comm_num)
#define Flash_Fail     3
#define Flash_Success  0
#endif
