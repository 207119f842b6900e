/* placeholder; filled below */
#include "mhm_oracle.h"
