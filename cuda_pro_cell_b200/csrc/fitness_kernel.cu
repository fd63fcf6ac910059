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
 * fixed order, so the result is reproducible run to run.
 */
#include "sim_kernels.h"

namespace procell_b200 {

constexpr int kFitThreads = 256;

__global__ void __launch_bounds__(kFitThreads) k_sweep_fitness(const long long* __restrict__ counts,   /* [S][K][T] */
                                                               const uint32_t* __restrict__ key_channel, /* [K], 0xFFFFFFFF = not a row */
                                                               const double* __restrict__ target_share, /* [C] */
                                                               uint32_t n_keys, uint32_t n_types, uint32_t n_channels,
                                                               double* __restrict__ out)                /* [S] */
{
    extern __shared__ unsigned long long s_acc[];      /* [n_channels] then kFitThreads doubles */
    double* s_red = reinterpret_cast<double*>(s_acc + n_channels);
    const uint32_t set = blockIdx.x;
    for (uint32_t c = threadIdx.x; c < n_channels; c += blockDim.x) s_acc[c] = 0ull;
    __syncthreads();
    const long long* base = counts + (size_t)set * n_keys * n_types;
    for (uint32_t key = threadIdx.x; key < n_keys; key += blockDim.x) {
        const uint32_t ch = key_channel[key];
        if (ch == 0xFFFFFFFFu) continue;
        unsigned long long sum = 0;
        for (uint32_t t = 0; t < n_types; ++t) sum += (unsigned long long)base[(size_t)key * n_types + t];
        if (sum) atomicAdd(&s_acc[ch], sum);
    }
    __syncthreads();
    /* total (exact integer), fixed-order tree reduction */
    unsigned long long part = 0;
    for (uint32_t c = threadIdx.x; c < n_channels; c += blockDim.x) part += s_acc[c];
    unsigned long long* s_tot = reinterpret_cast<unsigned long long*>(s_red);
    s_tot[threadIdx.x] = part;
    __syncthreads();
    for (int off = kFitThreads / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) s_tot[threadIdx.x] += s_tot[threadIdx.x + off];
        __syncthreads();
    }
    const double total = (double)s_tot[0];
    __syncthreads();
    double bc = 0.0;                                     /* Bhattacharyya coefficient, thread-strided then tree */
    if (total > 0.0)
        for (uint32_t c = threadIdx.x; c < n_channels; c += blockDim.x)
            bc += sqrt(((double)s_acc[c] / total) * target_share[c]);
    s_red[threadIdx.x] = bc;
    __syncthreads();
    for (int off = kFitThreads / 2; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) s_red[threadIdx.x] += s_red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double h2 = 1.0 - s_red[0];
        out[set] = total > 0.0 ? sqrt(h2 > 0.0 ? h2 : 0.0) : 1.0;
    }
}

cudaError_t launch_sweep_fitness(const long long* counts, const uint32_t* key_channel, const double* target_share,
                                 uint32_t n_sets, uint32_t n_keys, uint32_t n_types, uint32_t n_channels, double* out,
                                 cudaStream_t stream)
{
    const size_t smem = (size_t)n_channels * 8 + (size_t)kFitThreads * 8;
    cudaError_t e = cudaFuncSetAttribute(k_sweep_fitness, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_sweep_fitness<<<n_sets, kFitThreads, smem, stream>>>(counts, key_channel, target_share, n_keys, n_types, n_channels, out);
    return cudaGetLastError();
}

}  // namespace procell_b200
