// ABI version + device check entry points.
#include "common.cuh"
#include "../../include/heal_b200.h"

extern "C" int heal_abi_version(void) { return HEAL_B200_ABI_VERSION; }

extern "C" int heal_device_check(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return HEAL_ERR_DRIVER;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return HEAL_ERR_DRIVER;
    return (prop.major == 10) ? HEAL_OK : HEAL_ERR_UNSUPPORTED;
}
