// refsrc_driver.cpp — builds oracle/_ref/librefsrc.so: the REFERENCE'S OWN src/mcml_optim.cpp, src/mcml_full.cpp and src/mcml_la.cpp,
// compiled unmodified from where they lie under /root/reference/src, together with its own headers, against the stand-in headers of
// oracle/shim (TEST INFRASTRUCTURE; never shipped, never on the product path).
//
// The exported functions are those of tests/adapters_driver.cpp (drv_mcml_full, drv_mcml_optim, drv_mcml_hess, drv_aic_mcml, drv_mvn_ll,
// drv_mcmc_sample, drv_mcml_la): the SAME driver that calls this repo's Rcpp adapters (src/) calls the reference's own function bodies here,
// so the two libraries can be driven side by side with identical arguments (tests/test_rcpp_adapters.py, tests/test_oracle_vs_ref.py).
//
// What runs as the reference wrote it: the entry-point bodies (object construction, the MCML loop and its convergence test, cAIC, the
// Hessian set-up), every header of inst/include/glmmrmcml.  What is a stand-in: Eigen / Rcpp (oracle/shim/RcppEigen.h, Rcpp.h), glmmrBase
// (shim/glmmr.h: reconstruction), rminqa (shim/rbobyqa.h: optimhess / fmingr restated exactly, BOBYQA replaced by another bounded
// minimiser), SparseChol (stub: the *_sparse exports compile and throw).  Random numbers: as in ref_driver.cpp the R normals and the
// std::minstd_rand uniforms of mhmcmc.h:48-55,62,85 are redirected to the Philox stream of the device sampler — here for any number of
// sample() calls: the s-th call of a run (s = 1, 2, ...) is keyed by seed + s * 0x9E3779B97F4A7C15 when `golden` is set (what
// gmb_mcml_full does per MCML iteration) or by the seed itself (gmb_mcmc_sample).
#include <cstdint>
#include <cstring>
#include <random>
#include <string>

namespace refrng {
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
static inline double u01(uint32_t lo, uint32_t hi) { uint64_t x = ((uint64_t)hi << 32) | lo; return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static inline void uniform2(uint64_t seed, uint32_t idx, uint32_t iter, uint32_t chain, uint32_t stream, double* a, double* b) {
    uint32_t c[4] = {idx, iter, chain, stream};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    *a = u01(c[0], c[1]); *b = u01(c[2], c[3]);
}
static inline void normal_vec(uint64_t seed, uint32_t iter, uint32_t chain, uint32_t stream, int Q, double* z) {
    for (int p = 0; p < (Q + 1) / 2; p++) {
        double u1, u2; uniform2(seed, (uint32_t)p, iter, chain, stream, &u1, &u2);
        double r = std::sqrt(-2.0 * std::log(u1)), a = 2.0 * 3.14159265358979323846 * u2;
        z[2 * p] = r * std::cos(a);
        if (2 * p + 1 < Q) z[2 * p + 1] = r * std::sin(a);
    }
}
// proposals_per_sample = warmup + nsamp of the run's sample() calls (each call: initialise_u = 2 normal vectors, then one normal vector
// and one uniform per proposal); the constructor's initialise_u (mhmcmc.h:44) consumes the first two normal vectors of a run
struct State { uint64_t seed = 0; uint32_t chain = 0; long proposals_per_sample = 1; int golden = 0; long normal_calls = 0, uniform_calls = 0; };
static State g;
static inline uint64_t sample_seed(long s) { return g.golden ? g.seed + (uint64_t)(s + 1) * 0x9E3779B97F4A7C15ull : g.seed; }
static void normal_source(int n, double* out) {
    const long k = g.normal_calls++;
    const long per = g.proposals_per_sample + 2;
    if (k >= 2) {
        const long s = (k - 2) / per, w = (k - 2) % per;
        if (w == 0) { normal_vec(sample_seed(s), 0, g.chain, 0, n, out); return; }                      // sample()'s u_ (:127 -> :48)
        if (w >= 2) { normal_vec(sample_seed(s), (uint32_t)(w - 2), g.chain, 2, n, out); return; }     // proposal w - 2 (:62)
    }
    for (int i = 0; i < n; i++) out[i] = 0.0;                                                          // r_ of initialise_u: overwritten by the first proposal
}
static double next_uniform() {
    const long k = g.uniform_calls++;
    double a, b; uniform2(sample_seed(k / g.proposals_per_sample), 0, (uint32_t)(k % g.proposals_per_sample), g.chain, 3, &a, &b);
    return a;
}
}  // namespace refrng

// redirect the three <random> names used by mhmcmc.h:27-28,55-56,85
namespace std {
struct gmb_shim_rd { unsigned operator()() { return 0u; } };
struct gmb_shim_rng { gmb_shim_rng() {} explicit gmb_shim_rng(unsigned) {} };
template <class T> struct gmb_shim_dist { gmb_shim_dist() {} gmb_shim_dist(T, T) {} T operator()(gmb_shim_rng&) { return (T)refrng::next_uniform(); } };
}
#define random_device gmb_shim_rd
#define minstd_rand gmb_shim_rng
#define uniform_real_distribution gmb_shim_dist

#include <RcppEigen.h>
// the reference's own translation units (found through -I$(REFROOT)/src), one after the other in this one
#include <mcml_optim.cpp>
#include <mcml_full.cpp>
#include <mcml_la.cpp>

#undef random_device
#undef minstd_rand
#undef uniform_real_distribution

// the driver of tests/: plain buffers <-> the Rcpp-level signatures
#include "../tests/adapters_driver.cpp"

// Start of a run: the next entry-point call draws from this Philox stream (see the header comment).
extern "C" __attribute__((visibility("default")))
void refsrc_set_stream(unsigned long long seed, unsigned chain, long proposals_per_sample, int golden) {
    refrng::g = refrng::State();
    refrng::g.seed = seed; refrng::g.chain = chain; refrng::g.proposals_per_sample = proposals_per_sample > 0 ? proposals_per_sample : 1; refrng::g.golden = golden;
    Rcpp::normal_source() = refrng::normal_source;
}
extern "C" __attribute__((visibility("default"))) const char* refsrc_root() { return REF_ROOT; }
