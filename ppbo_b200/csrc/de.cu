// ppbo_b200 -- the sequential differential evolution behind GPModel.mu_star (src/gp_model.py:415-437), host side.
//
// The reference finds the maximiser of the posterior mean with scipy.optimize.differential_evolution(mu_pred_neq, bounds,
// updating='immediate', maxiter=2000): ~10^3..10^4 DEPENDENT evaluations of mu(x) per model update, every trial vector built from
// numpy's global legacy random stream.  With the evaluation on the GPU (ppbo_mu_pred_point) what is left of an evaluation is
// scipy's own per-trial Python work (30..45 us against ~12 us of launch + synchronise).  This file restates that loop in C++ so that
// the host work per trial is ~1 us, and it replays scipy 1.18 + numpy's legacy RandomState DRAW FOR DRAW: the same MT19937 words
// are consumed in the same order and every floating-point operation is the one numpy performs, so the trial vectors, the accepted
// members, the stopping generation and the state the stream is left in are bit for bit those of the scipy call
// (tests/test_de_replay.py compares them on the CPU for arbitrary objectives through a callback; tests/test_src_gpu.py on the GPU).
// The evaluations can be issued in speculative windows (Solver::generation_windows) without changing any of that.
//
// What is restated (scipy/optimize/_differentialevolution.py of scipy 1.18.1, defaults of the reference's call):
//   strategy 'best1bin', init 'latinhypercube', popsize 15 (population 15 D), mutation (0.5, 1) = dither per generation,
//   recombination 0.7, tol 0.01, atol 0, updating 'immediate', no constraints, no integrality.  The L-BFGS-B polish that follows
//   the evolution stays with scipy on the Python side (a few hundred evaluations through the same objective).
// numpy pieces (numpy/random/_mt19937, distributions.c, mtrand.pyx legacy paths; numpy/core pairwise summation):
//   random_sample = (a >> 5, b >> 6) / 2^53, uniform(l, h) = l + (h - l) u, randint by bit mask + rejection on 32-bit words,
//   shuffle = Fisher-Yates from the top with random_interval, add.reduce = first element + pairwise sum of the rest.
#include <cmath>
#include <limits>
#include <vector>

#include "../../include/ppbo_b200.h"
#include "common.cuh"

namespace ppbo {
namespace de {

struct MT19937 {                       // numpy/random/src/mt19937/mt19937.{h,c}
    uint32_t* key;
    int pos;
    void gen() {
        const uint32_t A = 0x9908b0dfu, UP = 0x80000000u, LO = 0x7fffffffu;
        int i;
        uint32_t y;
        for (i = 0; i < 624 - 397; ++i) {
            y = (key[i] & UP) | (key[i + 1] & LO);
            key[i] = key[i + 397] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        }
        for (; i < 623; ++i) {
            y = (key[i] & UP) | (key[i + 1] & LO);
            key[i] = key[i + (397 - 624)] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        }
        y = (key[623] & UP) | (key[0] & LO);
        key[623] = key[396] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        pos = 0;
    }
    uint32_t next32() {
        if (pos == 624) gen();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    double next_double() {             // mt19937_next_double
        const int32_t a = (int32_t)(next32() >> 5), b = (int32_t)(next32() >> 6);
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    double uniform(double low, double range) { return low + range * next_double(); }      // random_uniform
    // random_interval (shuffle) and the masked legacy randint share the rule: smallest bit mask >= max, reject above max
    uint32_t interval(uint32_t max) {
        if (max == 0) return 0;        // no word is drawn
        uint32_t mask = max;
        mask |= mask >> 1;
        mask |= mask >> 2;
        mask |= mask >> 4;
        mask |= mask >> 8;
        mask |= mask >> 16;
        uint32_t v;
        while ((v = (next32() & mask)) > max) {
        }
        return v;
    }
    void shuffle(int* a, int n) {      // RandomState._shuffle_raw
        for (int i = n - 1; i >= 1; --i) {
            const int j = (int)interval((uint32_t)i);
            const int t = a[j];
            a[j] = a[i];
            a[i] = t;
        }
    }
};

// numpy's pairwise summation (DOUBLE_pairwise_sum, unit stride)
static double pairwise_sum(const double* a, long n) {
    if (n < 8) {
        double res = 0.0;
        for (long i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        long i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    long n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
}
// np.add.reduce of a contiguous 1-D array: the first element seeds the accumulator, the rest is summed pairwise
static double add_reduce(const double* a, long n) { return n == 0 ? 0.0 : a[0] + pairwise_sum(a + 1, n - 1); }

struct Solver {
    ppbo_objective_fn f;
    ppbo_objective_batch_fn fb = nullptr;                  // evaluates B parameter vectors in one call (speculative windows)
    void* ctx;
    int D, n;
    std::vector<double> arg1, arg2, pop, energy, trial, bprime, params, tmp;
    std::vector<int> index;
    MT19937 rng;
    double scale = 0.0, recombination = 0.7;
    long nfev = 0;
    int failed = 0;
    bool own_error_text = true;        // false: the objective has already recorded why it failed

    double* row(int i) { return pop.data() + (size_t)i * D; }
    double evaluate(const double* t) {                     // _scale_parameters + func
        for (int d = 0; d < D; ++d) params[d] = arg1[d] + (t[d] - 0.5) * arg2[d];
        ++nfev;
        ++calls;
        const double e = f(params.data(), D, ctx);
        if (e != e && !failed) failed = (int)nfev;         // remember the first NaN (a failed device call reports itself this way)
        return e;
    }
    void init_lhs() {                                      // init_population_lhs
        const double segsize = 1.0 / n, step = 1.0 / n;
        std::vector<double> samples((size_t)n * D);
        for (int i = 0; i < n; ++i) {
            const double lin = (double)i * step;           // np.linspace(0., 1., n, endpoint=False)
            for (int d = 0; d < D; ++d) samples[(size_t)i * D + d] = segsize * rng.uniform(0.0, 1.0) + lin;
        }
        std::vector<int> order(n);
        for (int d = 0; d < D; ++d) {
            for (int i = 0; i < n; ++i) order[i] = i;
            rng.shuffle(order.data(), n);                  // rng.permutation(range(n))
            for (int i = 0; i < n; ++i) row(i)[d] = samples[(size_t)order[i] * D + d];
        }
    }
    void promote_lowest() {                                // _promote_lowest_energy: first arg-min (NaN wins, as in np.argmin)
        int l = 0;
        for (int i = 0; i < n; ++i) {
            if (energy[i] != energy[i]) {
                l = i;
                break;
            }
            if (energy[i] < energy[l]) l = i;
        }
        if (l == 0) return;
        const double e = energy[0];
        energy[0] = energy[l];
        energy[l] = e;
        for (int d = 0; d < D; ++d) {
            const double t = row(0)[d];
            row(0)[d] = row(l)[d];
            row(l)[d] = t;
        }
    }
    bool converged(double tol, double atol) {              // DifferentialEvolutionSolver.converged
        for (int i = 0; i < n; ++i)
            if (std::isinf(energy[i])) return false;
        const double mean = add_reduce(energy.data(), n) / (double)n;
        for (int i = 0; i < n; ++i) {
            const double x = energy[i] - mean;
            tmp[i] = x * x;
        }
        const double sd = std::sqrt(add_reduce(tmp.data(), n) / (double)n);
        return sd <= atol + tol * std::fabs(mean);
    }
    // trial vector of candidate c from the current population and the next draws of the stream: _mutate + _ensure_constraint
    int last_r0 = 0, last_r1 = 0;                          // the two sampled members the last make_trial read
    void make_trial(int c, double* out) {
        const int fill_point = (int)rng.interval((uint32_t)(D - 1));              // rng.randint(D)
        rng.shuffle(index.data(), n);                                              // _select_samples(c, 5)
        int r[2], k = 0;
        for (int i = 0; i < 6 && i < n && k < 2; ++i)
            if (index[i] != c) r[k++] = index[i];
        last_r0 = r[0];
        last_r1 = r[1];
        const double *p0 = row(0), *pa = row(r[0]), *pb = row(r[1]), *pc = row(c);
        for (int d = 0; d < D; ++d) {
            const double diff = pa[d] - pb[d];
            const double sdiff = scale * diff;
            bprime[d] = p0[d] + sdiff;                                             // _best1
        }
        for (int d = 0; d < D; ++d) {
            const bool cross = rng.uniform(0.0, 1.0) < recombination;
            out[d] = (cross || d == fill_point) ? bprime[d] : pc[d];
        }
        for (int d = 0; d < D; ++d)                                                // _ensure_constraint
            if (out[d] > 1.0 || out[d] < 0.0) out[d] = rng.uniform(0.0, 1.0);
    }
    // the compare-and-replace step of __next__: 0 rejected, 1 member c replaced, 2 replaced and the best-member branch taken
    int accept(int c, const double* t, double e) {
        if (!(e <= energy[c])) return 0;
        for (int d = 0; d < D; ++d) row(c)[d] = t[d];
        energy[c] = e;
        if (!(e <= energy[0])) return 1;
        promote_lowest();
        return 2;
    }
    void generation() {                                    // __next__, updating='immediate': one evaluation per call of f
        for (int c = 0; c < n; ++c) {
            make_trial(c, trial.data());
            accept(c, trial.data(), evaluate(trial.data()));
        }
    }
    // The same generation with the evaluations in windows.  A trial reads four members of the population: the best (row 0), two
    // sampled ones and its own; the draws it consumes do not depend on anything else.  So: build the next W trials from the
    // population as it is, evaluate them in ONE call and walk the results in order, exactly like the sequential loop.  A rejected
    // trial changes nothing.  An accepted trial replaces its own member only, unless it is also at least as good as the best: then
    // row 0 changes and every later trial of the window is stale.  Otherwise a later trial is stale only if one of its two sampled
    // members is a row replaced earlier in this window.  At the first stale trial the rest of the window is thrown away and the
    // stream and the shuffled index array are put back to where they were before that trial was built.  Every retained trial, draw
    // and comparison is the sequential loop's; only evaluations whose results are never looked at are added.
    std::vector<uint32_t> snap_key;
    std::vector<int> snap_pos, snap_index, win_r0, win_r1, replaced;
    std::vector<double> win_trial, win_params, win_energy;
    long wasted = 0, calls = 0;                            // discarded evaluations; calls of f / fb
    void generation_windows(int W) {
        if ((int)snap_pos.size() < W) {
            snap_key.resize((size_t)W * 624);
            snap_pos.resize(W);
            snap_index.resize((size_t)W * n);
            win_trial.resize((size_t)W * D);
            win_params.resize((size_t)W * D);
            win_energy.resize(W);
            win_r0.resize(W);
            win_r1.resize(W);
            replaced.reserve(W);
        }
        int c = 0;
        while (c < n) {
            const int B = n - c < W ? n - c : W;
            for (int k = 0; k < B; ++k) {
                memcpy(&snap_key[(size_t)k * 624], rng.key, sizeof(uint32_t) * 624);      // state BEFORE trial k is built
                snap_pos[k] = rng.pos;
                memcpy(&snap_index[(size_t)k * n], index.data(), sizeof(int) * n);
                double* t = &win_trial[(size_t)k * D];
                make_trial(c + k, t);
                win_r0[k] = last_r0;
                win_r1[k] = last_r1;
                for (int d = 0; d < D; ++d) win_params[(size_t)k * D + d] = arg1[d] + (t[d] - 0.5) * arg2[d];
            }
            ++calls;
            if (fb(win_params.data(), B, D, win_energy.data(), ctx) != 0) {
                if (!failed) failed = (int)nfev + 1;
                return;
            }
            int k = 0;
            replaced.clear();
            while (k < B) {
                bool stale = false;
                for (int r : replaced) stale = stale || r == win_r0[k] || r == win_r1[k];
                if (stale) break;
                const double e = win_energy[k];
                ++nfev;
                if (e != e && !failed) failed = (int)nfev;
                const int a = accept(c + k, &win_trial[(size_t)k * D], e);
                if (a == 1) replaced.push_back(c + k);
                ++k;
                if (a == 2) break;
            }
            if (k < B) {                                   // trials k .. B-1 were built from a population that no longer exists
                memcpy(rng.key, &snap_key[(size_t)k * 624], sizeof(uint32_t) * 624);
                rng.pos = snap_pos[k];
                memcpy(index.data(), &snap_index[(size_t)k * n], sizeof(int) * n);
                wasted += B - k;
            }
            c += k;
            if (failed) return;
        }
    }
};

static int run(ppbo_objective_fn f, ppbo_objective_batch_fn fb, int window, void* ctx, int D, const double* lower, const double* upper,
               int popsize, int maxiter, double tol, double atol, double mutation_lo, double mutation_hi, double recombination,
               unsigned int* mt_key, int* mt_pos, double* x_out, double* fun_out, int* stats, bool own_error_text = true) {
    PPBO_REQUIRE(f != nullptr && D >= 1 && D <= PPBO_MAX_D && popsize >= 1 && maxiter >= 0, "problem");
    PPBO_REQUIRE(window >= 1 && window <= PPBO_MAX_POINTS, "window");
    if (fb == nullptr) window = 1;
    PPBO_REQUIRE(lower != nullptr && upper != nullptr && mt_key != nullptr && mt_pos != nullptr && x_out != nullptr && fun_out != nullptr,
                 "null pointer");
    PPBO_REQUIRE(*mt_pos >= 0 && *mt_pos <= 624, "position in the MT19937 state");
    PPBO_REQUIRE(mutation_lo <= mutation_hi && mutation_lo >= 0.0 && mutation_hi < 2.0, "mutation range");
    PPBO_REQUIRE(recombination >= 0.0 && recombination <= 1.0, "recombination");
    Solver s;
    s.f = f;
    s.fb = fb;
    s.ctx = ctx;
    s.D = D;
    int varying = 0;
    for (int d = 0; d < D; ++d) {
        PPBO_REQUIRE(std::isfinite(lower[d]) && std::isfinite(upper[d]), "bounds must be finite");
        varying += lower[d] != upper[d];
    }
    s.n = popsize * (varying > 1 ? varying : 1);
    if (s.n < 5) s.n = 5;
    const int n = s.n;
    s.arg1.resize(D);
    s.arg2.resize(D);
    for (int d = 0; d < D; ++d) {
        s.arg1[d] = 0.5 * (lower[d] + upper[d]);
        s.arg2[d] = std::fabs(lower[d] - upper[d]);
    }
    s.pop.assign((size_t)n * D, 0.0);
    s.energy.assign(n, std::numeric_limits<double>::infinity());
    s.trial.resize(D);
    s.bprime.resize(D);
    s.params.resize(D);
    s.tmp.resize(n);
    s.index.resize(n);
    for (int i = 0; i < n; ++i) s.index[i] = i;
    s.rng.key = mt_key;
    s.rng.pos = *mt_pos;
    s.recombination = recombination;
    s.own_error_text = own_error_text;

    s.init_lhs();
    if (window > 1) {                                                       // solve(): initial energies, in population order
        std::vector<double> par((size_t)window * D);
        for (int i0 = 0; i0 < n && !s.failed; i0 += window) {
            const int B = n - i0 < window ? n - i0 : window;
            for (int k = 0; k < B; ++k)
                for (int d = 0; d < D; ++d) par[(size_t)k * D + d] = s.arg1[d] + (s.row(i0 + k)[d] - 0.5) * s.arg2[d];
            ++s.calls;
            if (fb(par.data(), B, D, &s.energy[i0], ctx) != 0) s.failed = (int)s.nfev + 1;
            for (int k = 0; k < B && !s.failed; ++k) {
                ++s.nfev;
                if (s.energy[i0 + k] != s.energy[i0 + k]) s.failed = (int)s.nfev;
            }
        }
    } else {
        for (int i = 0; i < n; ++i) s.energy[i] = s.evaluate(s.row(i));
    }
    s.promote_lowest();
    int nit = 0, conv = 0;
    for (nit = 1; nit <= maxiter; ++nit) {
        s.scale = s.rng.uniform(mutation_lo, mutation_hi - mutation_lo);    // dither
        if (window > 1) s.generation_windows(window);
        else s.generation();
        if (s.failed) break;
        if (s.converged(tol, atol)) {
            conv = 1;
            break;
        }
    }
    if (nit > maxiter) nit = maxiter;                                        // range(1, maxiter + 1) ran out
    *mt_pos = s.rng.pos;
    for (int d = 0; d < D; ++d) x_out[d] = s.arg1[d] + (s.row(0)[d] - 0.5) * s.arg2[d];
    *fun_out = s.energy[0];
    if (stats) {
        stats[0] = nit;
        stats[1] = (int)s.nfev;
        stats[2] = conv;
        stats[3] = n;
        stats[4] = (int)s.wasted;
        stats[5] = (int)s.calls;
    }
    if (s.failed) {
        if (s.own_error_text) set_error("the objective returned NaN at evaluation %d", s.failed);
        return PPBO_ERR_ARG;
    }
    return PPBO_OK;
}

struct MuCtx {
    int kind, N, D;
    const double *X, *ls, *alpha;
    double sigma_f;
    void* stream;
    int rc;
};
static double neg_mu(const double* x, int D, void* p) {
    MuCtx* c = static_cast<MuCtx*>(p);
    double mu = 0.0;
    const int rc = ppbo_mu_pred_point(c->kind, c->X, c->N, D, c->ls, c->sigma_f, c->alpha, x, &mu, c->stream);
    if (rc) {
        c->rc = rc;
        return std::numeric_limits<double>::quiet_NaN();
    }
    return -mu;
}

static int neg_mu_batch(const double* x, int B, int D, double* out, void* p) {
    MuCtx* c = static_cast<MuCtx*>(p);
    const int rc = ppbo_mu_pred_points(c->kind, c->X, c->N, D, c->ls, c->sigma_f, c->alpha, x, B, out, c->stream);
    if (rc) {
        c->rc = rc;
        return rc;
    }
    for (int k = 0; k < B; ++k) out[k] = -out[k];
    return 0;
}

}  // namespace de
}  // namespace ppbo

extern "C" int ppbo_de_minimize(ppbo_objective_fn f, ppbo_objective_batch_fn f_batch, int window, void* ctx, int D,
                                const double* lower_h, const double* upper_h, int popsize, int maxiter, double tol, double atol,
                                double mutation_lo, double mutation_hi, double recombination, unsigned int* mt_key, int* mt_pos,
                                double* x_h, double* fun_h, int* stats_h) {
    return ppbo::de::run(f, f_batch, window, ctx, D, lower_h, upper_h, popsize, maxiter, tol, atol, mutation_lo, mutation_hi,
                         recombination, mt_key, mt_pos, x_h, fun_h, stats_h);
}

extern "C" int ppbo_mu_star_de(int kind, const double* X, int N, int D, const double* lengthscales_h, double sigma_f,
                               const double* alpha, const double* lower_h, const double* upper_h, int popsize, int maxiter, double tol,
                               double atol, double mutation_lo, double mutation_hi, double recombination, int window,
                               unsigned int* mt_key, int* mt_pos, double* x_h, double* fun_h, int* stats_h, void* stream) {
    ppbo::de::MuCtx c{kind, N, D, X, lengthscales_h, alpha, sigma_f, stream, 0};
    const int rc = ppbo::de::run(ppbo::de::neg_mu, ppbo::de::neg_mu_batch, window, &c, D, lower_h, upper_h, popsize, maxiter, tol, atol,
                                 mutation_lo, mutation_hi, recombination, mt_key, mt_pos, x_h, fun_h, stats_h, false);
    if (rc && !c.rc) ppbo::set_error("the posterior mean is NaN inside the differential evolution");
    return c.rc ? c.rc : rc;            // a failed device call keeps its own status and error text
}
