#include "launch.cuh"
namespace ofdm {
OFDM_DEFINE_KERNELSET(12)
}
