/* fitness_kernel.cu - on-GPU fitness of a parameter sweep ("next" row 1).
 *
 * A calibration sweep (BASELINE config 5) produces one histogram per parameter set: 1024 x 4420 x 3 int64 = 104 MB
 * that a fitting loop would have to pull over PCIe only to reduce each to one number.  This kernel does the reduction
 * where the counts are: per set, re-bin the simulated rows onto the target histogram's channels and compute the
 * Hellinger distance to the target.
 *
 * Re-binning follows the reference's (dead, never called) helper utils::rebin, src/utils/util.cu:111-138:
 * a value goes to the first channel whose value is >= it (values beyond the last channel go to the last one).
 * The key -> channel map is built on the host (capi.cu: procell_engine_set_target); distance:
 *     H = sqrt(1 - sum_c sqrt(p_c * q_c)),   p = simulated channel shares, q = target channel shares.
 * One CTA per parameter set; channel sums are exact integers in shared memory, the floating-point reduction runs in a
 * fixed order, so the result is reproducible run to run.  The arithmetic lives in fitness_device.h, which the simulation
 * kernel's own tail uses too (fitness in the same launch); this kernel serves count tensors that did not come from the
 * engine's last launch (a reduced multi-GPU tensor, a target set after the run).
 */
#include "sim_kernels.h"
#include "fitness_device.h"

namespace procell_b200 {

__global__ void __launch_bounds__(kFitThreads) k_sweep_fitness(const long long* __restrict__ counts,   /* [S][K][T] */
                                                               const uint32_t* __restrict__ key_channel, /* [K], 0xFFFFFFFF = not a row */
                                                               const double* __restrict__ target_share, /* [C] */
                                                               uint32_t n_keys, uint32_t n_types, uint32_t n_channels,
                                                               double* __restrict__ out)                /* [S] */
{
    extern __shared__ unsigned long long s_acc[];      /* [n_channels] then kFitThreads doubles */
    const uint32_t set = blockIdx.x;
    fitness_of_set<false>(counts + (size_t)set * n_keys * n_types, key_channel, target_share, n_keys, n_types, n_channels, s_acc, out + set);
}

cudaError_t launch_sweep_fitness(const long long* counts, const uint32_t* key_channel, const double* target_share,
                                 uint32_t n_sets, uint32_t n_keys, uint32_t n_types, uint32_t n_channels, double* out,
                                 cudaStream_t stream)
{
    const size_t smem = fitness_smem_bytes(n_channels);
    cudaError_t e = cudaFuncSetAttribute(k_sweep_fitness, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_sweep_fitness<<<n_sets, kFitThreads, smem, stream>>>(counts, key_channel, target_share, n_keys, n_types, n_channels, out);
    return cudaGetLastError();
}

}  // namespace procell_b200
