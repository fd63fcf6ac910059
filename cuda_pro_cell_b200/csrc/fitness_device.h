/* fitness_device.h - the per-parameter-set fitness of a calibration sweep as a device function, shared by the
 * stand-alone kernel (fitness_kernel.cu: k_sweep_fitness, one CTA per set) and by the tail of the simulation kernel
 * itself (sim_kernels.cu: after the last CTA has flushed its table, the CTAs of the SAME launch share the sets among
 * them and reduce each slab while it still sits in L2 - SURVEY.md section 8f, row 1: "in the same launch").
 *
 * Re-binning follows the reference's (dead, never called) helper utils::rebin, src/utils/util.cu:111-138: a value goes
 * to the first channel whose value is >= it (values beyond the last channel go to the last one); the key -> channel
 * map is built on the host (capi.cu: procell_engine_set_target).  Distance:
 *     H = sqrt(1 - sum_c sqrt(p_c * q_c)),   p = simulated channel shares, q = target channel shares.
 * Channel sums are exact integers (any number of threads may add to them); the two floating-point reductions run over
 * the first kFitThreads threads in a fixed order, so the result does not depend on the size of the calling CTA and is
 * the same, bit for bit, from either caller.
 */
#ifndef PROCELL_FITNESS_DEVICE_H
#define PROCELL_FITNESS_DEVICE_H

#include <stdint.h>

namespace procell_b200 {

constexpr int kFitThreads = 256;

/* shared memory the caller provides: n_channels u64 accumulators followed by kFitThreads doubles */
__host__ __device__ constexpr size_t fitness_smem_bytes(uint32_t n_channels) { return (size_t)n_channels * 8 + (size_t)kFitThreads * 8; }

/* called by ALL threads of the CTA (blockDim.x >= kFitThreads); counts_set = the [n_keys][n_types] slab of one set.
 * FROM_L2: the slab was written by other SMs' atomics in this very launch - read it with ld.global.cg */
template <bool FROM_L2>
__device__ __forceinline__ void fitness_of_set(const long long* __restrict__ counts_set, const uint32_t* __restrict__ key_channel,
                                               const double* __restrict__ target_share, uint32_t n_keys, uint32_t n_types,
                                               uint32_t n_channels, unsigned long long* s_acc, double* out)
{
    double* s_red = reinterpret_cast<double*>(s_acc + n_channels);
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    for (uint32_t c = tid; c < n_channels; c += nthr) s_acc[c] = 0ull;
    __syncthreads();
    for (uint32_t key = tid; key < n_keys; key += nthr) {
        const uint32_t ch = __ldg(key_channel + key);
        if (ch == 0xFFFFFFFFu) continue;
        unsigned long long sum = 0;
        for (uint32_t t = 0; t < n_types; ++t) {
            const long long* p = counts_set + (size_t)key * n_types + t;
            sum += (unsigned long long)(FROM_L2 ? __ldcg(p) : *p);
        }
        if (sum) atomicAdd(&s_acc[ch], sum);
    }
    __syncthreads();
    /* total (exact integer), fixed-order tree reduction over kFitThreads threads */
    unsigned long long* s_tot = reinterpret_cast<unsigned long long*>(s_red);
    if (tid < (uint32_t)kFitThreads) {
        unsigned long long part = 0;
        for (uint32_t c = tid; c < n_channels; c += kFitThreads) part += s_acc[c];
        s_tot[tid] = part;
    }
    __syncthreads();
    for (int off = kFitThreads / 2; off > 0; off >>= 1) {
        if ((int)tid < off) s_tot[tid] += s_tot[tid + off];
        __syncthreads();
    }
    const double total = (double)s_tot[0];
    __syncthreads();
    if (tid < (uint32_t)kFitThreads) {
        double bc = 0.0;                                 /* Bhattacharyya coefficient, thread-strided then tree */
        if (total > 0.0)
            for (uint32_t c = tid; c < n_channels; c += kFitThreads)
                bc += sqrt(((double)s_acc[c] / total) * __ldg(target_share + c));
        s_red[tid] = bc;
    }
    __syncthreads();
    for (int off = kFitThreads / 2; off > 0; off >>= 1) {
        if ((int)tid < off) s_red[tid] += s_red[tid + off];
        __syncthreads();
    }
    if (tid == 0) {
        const double h2 = 1.0 - s_red[0];
        *out = total > 0.0 ? sqrt(h2 > 0.0 ? h2 : 0.0) : 1.0;
    }
    __syncthreads();
}

}  // namespace procell_b200
#endif
