/* drop-in forwarder: see sr_compat.h */
#include "sr_compat.h"
