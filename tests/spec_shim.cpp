/* spec_shim.cpp - TEST-ONLY host build of the product's arithmetic header (cuda_pro_cell_b200/csrc/procell_spec.h).
 * The product never runs this arithmetic on the CPU; the shim lets the CPU test-suite check, without a GPU, that
 * the header the kernels include states the same operation sequence as the oracle (bit-for-bit). */
#include "../cuda_pro_cell_b200/csrc/procell_spec.h"

static const struct {
    uint64_t log_rows[1 << PCM_LOG_N_BITS][2];
    uint64_t sincos_rows[1 << PCM_SC_N_BITS][2];
    uint64_t zig_rows[1 << PCM_ZIG_N_BITS][2];
    uint64_t zig_wedge_rows[1 << PCM_ZIG_N_BITS][2];
} kTab = { { PCM_LOG_TABLE_ROWS }, { PCM_SINCOS_TABLE_ROWS }, { PCM_ZIG_TABLE_ROWS }, { PCM_ZIG_WEDGE_ROWS } };
static_assert(sizeof(kTab) == PCS_TAB_ALL_DOUBLES * 8, "math table layout");
static const double* const kRows = reinterpret_cast<const double*>(&kTab);

extern "C" {
void shim_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out)
{
    pcs_u32x4 o = pcs_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
}
void shim_draw(uint32_t root, uint32_t set, uint32_t retry, uint32_t tag, uint64_t heap, uint64_t seed, uint32_t* out)
{
    pcs_u32x4 o = pcs_draw(root, set, retry, tag, heap, (uint32_t)seed, (uint32_t)(seed >> 32));
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
}
double shim_u53(uint32_t lo, uint32_t hi) { return pcs_u53(lo, hi); }
double shim_neg2log(double u) { return pcs_neg2log(u, kRows); }
void shim_sincos2pi(uint64_t v, double* s, double* c) { pcs_sincos2pi(v, kRows + PCS_TAB_SINCOS, s, c); }
void shim_normal_pair(const uint32_t* w, double u_forced, double* z)
{
    pcs_u32x4 b; b.x = w[0]; b.y = w[1]; b.z = w[2]; b.w = w[3];
    pcs_normal_pair(b, kRows, u_forced, &z[0], &z[1]);
}
/* one whole ziggurat trial for daughter c of the division whose block (tag 0, this retry) is w: 1 = accepted */
int shim_zig_trial(const uint32_t* w, uint32_t c, uint32_t root, uint32_t set, uint32_t retry, uint64_t heap, uint64_t seed, double* z)
{
    uint32_t rk[20];
    pcs_round_keys((uint32_t)seed, (uint32_t)(seed >> 32), rk);
    pcs_u32x4 b; b.x = w[0]; b.y = w[1]; b.z = w[2]; b.w = w[3];
    return pcs_zig_trial(b, c, z, root, set, retry, heap, rk, kRows, kRows + PCS_TAB_WEDGE) ? 1 : 0;
}
/* the fast test alone (what the common DIVIDE iteration evaluates): 1 = accepted; *z is set either way */
int shim_zig_fast(uint32_t lo, uint32_t hi, double* z) { return pcs_zig_fast(lo, hi, kRows + PCS_TAB_ZIG, z) ? 1 : 0; }
/* The form the KERNELS evaluate (sim_kernels.cu: zig_fast_smem / zig_fast_signed): the sign goes on the multiplicand
 * instead of being OR-ed onto the product, the test is |x| < x_next.  Restated here operation for operation so that the
 * CPU suite can show it is pcs_zig_fast bit for bit - value (a signed zero included) and verdict - on every layer. */
static int zig_fast_kernel_form(uint32_t lo, uint32_t hi, double* z)
{
    const double* row = kRows + PCS_TAB_ZIG + 2u * PCS_ZIG_LAYER(hi);
    const double x = PCS_MUL(pcs_bits2d(((uint64_t)(hi & 0x801FFFFFu) << 32) | (uint64_t)lo), row[0]);
    *z = x;
    return (x < 0.0 ? -x : x) < row[1] ? 1 : 0;
}
int shim_zig_fast_kernel_form(uint32_t lo, uint32_t hi, double* z) { return zig_fast_kernel_form(lo, hi, z); }
/* n pseudo-random draws (xorshift64*, every layer and both signs come up) + the edge mantissas of every layer: number of
 * draws on which the two forms differ in verdict or in the BITS of z */
uint64_t shim_zig_fast_forms_differ(uint64_t n, uint64_t state)
{
    uint64_t bad = 0;
    for (uint64_t i = 0; i < n; ++i) {
        state ^= state >> 12; state ^= state << 25; state ^= state >> 27;
        const uint64_t r = state * 0x2545F4914F6CDD1DULL;
        double a, b;
        const int fa = pcs_zig_fast((uint32_t)r, (uint32_t)(r >> 32), kRows + PCS_TAB_ZIG, &a) ? 1 : 0;
        const int fb = zig_fast_kernel_form((uint32_t)r, (uint32_t)(r >> 32), &b);
        bad += (fa != fb) || pcs_d2bits(a) != pcs_d2bits(b);
    }
    static const uint64_t edge[] = { 0, 1, 2, 0xFFFFFFFFull, 0x100000000ull, 0x000FFFFFFFFFFFFFull, 0x0010000000000000ull,
                                     0x001FFFFFFFFFFFFEull, 0x001FFFFFFFFFFFFFull };
    for (uint32_t layer = 0; layer < (1u << PCM_ZIG_N_BITS); ++layer)
        for (uint32_t sign = 0; sign < 2; ++sign)
            for (uint64_t m : edge) {
                const uint32_t hi = (sign << 31) | (layer << (31 - PCM_ZIG_N_BITS)) | (uint32_t)(m >> 32), lo = (uint32_t)m;
                double a, b;
                const int fa = pcs_zig_fast(lo, hi, kRows + PCS_TAB_ZIG, &a) ? 1 : 0;
                const int fb = zig_fast_kernel_form(lo, hi, &b);
                bad += (fa != fb) || pcs_d2bits(a) != pcs_d2bits(b);
            }
    return bad;
}
double shim_timer(double mean, double sd, double z) { return pcs_timer(mean, sd, z); }
double shim_u32unit(uint32_t m) { return pcs_u32unit(m); }
double shim_seed_normal(const uint32_t* w, double u_forced)
{
    pcs_u32x4 b; b.x = w[0]; b.y = w[1]; b.z = w[2]; b.w = w[3];
    return pcs_seed_normal(b, kRows, u_forced);
}
}
