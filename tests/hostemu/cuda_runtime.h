/* <cuda_runtime.h> for translation units run by the CPU emulator — TEST INFRASTRUCTURE ONLY */
#include "hostemu.h"
