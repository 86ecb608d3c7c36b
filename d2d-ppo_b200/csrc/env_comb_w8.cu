// CombinatorialEnv step kernels for 32-byte packet-buffer records (max deadline <= 32).
#include "env_comb_step.cuh"
namespace d2d {
template int launch_comb_step<8>(const StepArgs&, int, int, int, cudaStream_t);
}
