// TEST INFRASTRUCTURE ONLY -- CPU emulator of the pdeop backend.
//
// Implements pdeop_backend.h with plain sequential loops over the SAME per-element bodies
// (csrc/pdeop_elem.h) the CUDA kernels call, so that the index algebra, the axis tables, the
// wavefront Gauss-Seidel ordering and the host-side orchestration (V-cycle, FGMRES, gradients) can be
// checked against the oracle in a container without a GPU.  It is built into
// tests/emu/libpdeop_emu.so by tests/emu/build.py and loaded only by tests.  The product package
// never loads it: mech_nn_discovery_pde_b200 raises if the CUDA library or a GPU is missing.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "../../mech_nn_discovery_pde_b200/csrc/pdeop_backend.h"
#include "../../mech_nn_discovery_pde_b200/csrc/pdeop_elem.h"
#include "../../mech_nn_discovery_pde_b200/csrc/pdeop_gs_line.h"

#include <random>

namespace pdeop {

static int g_gs_reverse = 0;  // process pipelined sweeps in reverse order inside a step (order-independence check)
extern "C" void pdeop_emu_set_gs_reverse(int v) { g_gs_reverse = v; }

const char* be_name() { return "host-emulator"; }
void* be_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void be_free(void* p) { free(p); }
void be_upload(void* dst, const void* src, size_t bytes) { if (bytes) memcpy(dst, src, bytes); }
void be_zero(stream_t, void* p, size_t bytes) { memset(p, 0, bytes); }
int be_last_error(char*, int) { return 0; }
void* be_event_create() { return nullptr; }
void be_event_destroy(void*) {}
void be_event_record(void*, stream_t) {}
float be_event_elapsed_ms(void*, void*) { return 0.f; }
long long be_launch_count() { return 0; }
int be_current_device() { return -1; }
int be_default_gs_pipe() { return 2; }
bool be_default_chain() { return false; }


// ---- converged mode (plain loops; same semantics as the CUDA kernels) ----
void be_copy(stream_t, void* dst, const void* src, size_t bytes) { if (bytes) memcpy(dst, src, bytes); }
void be_bdot(stream_t, size_t n, int B, const double* a, const double* c, double* out, const int* done) {
    if (done && *done) return;
    for (int b = 0; b < B; ++b) {
        double acc = 0.0;
        for (size_t i = 0; i < n; ++i) acc += a[b * n + i] * c[b * n + i];
        out[b] += acc;
    }
}
static double safe_ratio(double a, double b) {
    const double q = a / b;
    return std::isfinite(q) ? q : 0.0;
}
void be_pcg_xr(stream_t, size_t n, int B, double* x, double* r, const double* p, const double* Ap, const double* rz,
               const double* pAp, const double* active, double* rr, const int* done) {
    if (done && *done) return;
    for (int b = 0; b < B; ++b) {
        const double alpha = active[b] != 0.0 ? safe_ratio(rz[b], pAp[b]) : 0.0;
        double acc = 0.0;
        for (size_t i = 0; i < n; ++i) {
            x[b * n + i] += alpha * p[b * n + i];
            r[b * n + i] -= alpha * Ap[b * n + i];
            acc += r[b * n + i] * r[b * n + i];
        }
        rr[b] += acc;
    }
}
void be_pcg_p(stream_t, size_t n, int B, double* p, const double* z, const double* rz_new, const double* rz,
              const double* active, const int* done) {
    if (done && *done) return;
    for (int b = 0; b < B; ++b) {
        const double beta = active[b] != 0.0 ? safe_ratio(rz_new[b], rz[b]) : 0.0;
        for (size_t i = 0; i < n; ++i) p[b * n + i] = z[b * n + i] + beta * p[b * n + i];
    }
}
void be_pcg_scalars(stream_t, int B, double* active, double* rz, double* rz_new, double* pAp, double* rr, double* bnorm,
                    double rtol, FgmresState* s, int first) {
    if (!first && s->done) return;
    int any = 0;
    double worst = 0.0;
    for (int b = 0; b < B; ++b) {
        if (first) {
            bnorm[b] = sqrt(rr[b]);
            active[b] = bnorm[b] > 0.0 ? 1.0 : 0.0;
            rz[b] = 0.0;
        } else {
            const double rel = bnorm[b] > 0.0 ? sqrt(rr[b]) / bnorm[b] : 0.0;
            if (!(rel > rtol)) active[b] = 0.0;
            worst = std::max(worst, rel);
            rz[b] = rz_new[b];
        }
        rz_new[b] = pAp[b] = rr[b] = 0.0;
        if (active[b] != 0.0) any = 1;
    }
    if (first) { s->iters = 0; s->rnorm = 0.0; } else { s->iters += 1; s->rnorm = worst; }
    s->done = any ? 0 : 1;
}
void be_poly_update(stream_t, size_t n, int B, double* x, double* d, const double* r, const double* dinv, double c1,
                    double c2, const double* lam, int mode, const int* done) {
    if (done && *done) return;
    for (int b = 0; b < B; ++b) {
        double c2b = c2;
        if (mode == 1) c2b = c2 / lam[b];
        else if (mode == 2) c2b = std::min(c2, 1.8 / lam[b]);
        for (size_t i = 0; i < n; ++i) {
            const double dn = c1 * d[b * n + i] + c2b * dinv[b * n + i] * r[b * n + i];
            d[b * n + i] = dn;
            x[b * n + i] += dn;
        }
    }
}
void be_restrict_t(stream_t, const LevelDev& Lf, const LevelDev& Lc, int B, int C, const double* in, double* out,
                   const int* done) {
    if (done && *done) return;
    memset(out, 0, (size_t)B * C * Lc.G * sizeof(double));
    for (int b = 0; b < B; ++b)
        for (int w = 0; w < Lf.G; ++w)
            restrict_t_elem(Lf, Lc, C, in + (size_t)b * C * Lf.G, out + (size_t)b * C * Lc.G, w);
}
void be_power_step(stream_t, size_t n, int B, double* v, const double* Kv, const double* dinv, double* lam, double*) {
    for (int b = 0; b < B; ++b) {
        double nn = 0.0;
        for (size_t i = 0; i < n; ++i) {
            v[b * n + i] = dinv[b * n + i] * Kv[b * n + i];
            nn += v[b * n + i] * v[b * n + i];
        }
        nn = sqrt(nn);
        lam[b] = nn;
        const double inv = nn > 0.0 ? 1.0 / nn : 0.0;
        for (size_t i = 0; i < n; ++i) v[b * n + i] *= inv;
    }
}

static inline size_t vstride(const LevelDev& L) { return (size_t)L.M * L.G; }
static inline size_t tstride(const LevelDev& L) { return (size_t)L.D * kTabEntries * kTabPitch; }

void be_build_tables(stream_t, const LevelDev& L, int B, const double* cv, const double* fv, const double* bv,
                     double* T) {
    for (int b = 0; b < B; ++b)
        for (int a = 0; a < L.D; ++a)
            for (int ip = 0; ip < L.P; ++ip)
                build_table_elem(L, a, ip, cv + (size_t)b * L.Ntot * 12, fv + (size_t)b * L.Ftot * 4,
                                 bv + (size_t)b * L.Ftot * 4, T + b * tstride(L) + (size_t)a * kTabEntries * kTabPitch);
}

void be_pack(stream_t, const LevelDev& L, int B, const double* api, double* wave) {
    for (int b = 0; b < B; ++b)
        for (int w = 0; w < L.G; ++w) pack_elem(L, api + b * vstride(L), wave + b * vstride(L), w);
}

void be_unpack(stream_t, const LevelDev& L, int B, const double* wave, double* api) {
    for (int b = 0; b < B; ++b)
        for (int w = 0; w < L.G; ++w) unpack_elem(L, wave + b * vstride(L), api + b * vstride(L), w);
}

void be_interp(stream_t, const LevelDev& Li, const LevelDev& Lo, int B, int C, const double* in, double* out, int add,
               const int* done) {
    if (done && *done) return;
    for (int b = 0; b < B; ++b)
        for (int w = 0; w < Lo.G; ++w)
            interp_elem(Li, Lo, C, in + (size_t)b * C * Li.G, out + (size_t)b * C * Lo.G, w, add);
}

void be_atb(stream_t, const LevelDev& L, int B, const double* coef, const double* rhs_nat, const double* iv_rhs,
            double* atb) {
    for (int b = 0; b < B; ++b) {
        for (int w = 0; w < L.G; ++w) atb_elem(L, coef + b * vstride(L), rhs_nat + (size_t)b * L.G, atb + b * vstride(L), w);
        for (int k = 0; k < L.n_init; ++k) atb_init_elem(L, iv_rhs + (size_t)b * L.n_init, atb + b * vstride(L), k);
    }
}

template <int D>
static void apply_k_t(const LevelDev& L, int B, const double* T, const double* coef, const double* x, const double* b,
                      double* y, int mode) {
    for (int ib = 0; ib < B; ++ib)
        for (int w = 0; w < L.G; ++w)
            apply_k_elem<D>(L, T + ib * tstride(L), coef + ib * vstride(L), x + ib * vstride(L),
                            b ? b + ib * vstride(L) : nullptr, y + ib * vstride(L), w, mode);
}

void be_apply_k(stream_t, const LevelDev& L, int B, const double* T, const double* coef, const double* x,
                const double* b, double* y, int mode, const int* done) {
    if (done && *done) return;
    if (L.D == 1) apply_k_t<1>(L, B, T, coef, x, b, y, mode);
    else if (L.D == 2) apply_k_t<2>(L, B, T, coef, x, b, y, mode);
    else apply_k_t<3>(L, B, T, coef, x, b, y, mode);
}

template <int D>
static void gs_t(const LevelDev& L, int B, const double* T, const double* coef, const double* dinv, const double* b,
                 double* x, int nsweeps) {
    std::vector<int> hs(L.S + 1);
    memcpy(hs.data(), L.hstart, sizeof(int) * (L.S + 1));
    const int steps = L.S + kGsLag * (nsweeps - 1);
    for (int ib = 0; ib < B; ++ib)
        for (int t = 0; t < steps; ++t)
            for (int kk = 0; kk < nsweeps; ++kk) {
                const int k = g_gs_reverse ? nsweeps - 1 - kk : kk;
                const int s = t - kGsLag * k;
                if (s < 0 || s >= L.S) continue;
                for (int w = hs[s]; w < hs[s + 1]; ++w)
                    gs_elem<D, LdPlain, kTabPitch>(L, L.rowbase, T + ib * tstride(L), coef + ib * vstride(L),
                                                   dinv + ib * vstride(L), b + ib * vstride(L), x + ib * vstride(L), w);
            }
}

// ---- line-marching Gauss-Seidel (csrc/pdeop_gs_line.h) under a discrete-event model of the CUDA kernel ----
// Every thread of every CTA of an instance is a little state machine over the phases A(0) B(0) A(1) B(1) ...; the
// scheduler runs one phase of a randomly chosen runnable thread at a time, under exactly the ordering the kernel's
// synchronisation provides and nothing more:
//   * split CTA barrier: A(n+1) of a thread may run once every thread of its CTA has finished A(n);
//   * cp.async: a copy issued in B(n) lands at a random moment between its issue and the end of the issuing thread's
//     A(n+1) (where the kernel waits for it);
//   * progress counters: thread 0 of a CTA publishes n at the start of its B(n) and may only continue once the
//     neighbouring CTAs have published what B(n+2) needs (same formulas as the kernel).
// A protocol error shows up as a result that differs from the sequential sweep, or as a deadlock (returns nonzero).
static unsigned g_line_seed = 12345;
static int g_line_maxthreads = 0;
extern "C" void pdeop_emu_set_line_seed(unsigned v) { g_line_seed = v; }
// test hook: fewer threads per CTA than the kernel uses, so that small grids span several CTAs
extern "C" void pdeop_emu_set_line_max_threads(int v) { g_line_maxthreads = v; }
static int g_line_last = 0;      // 1: the last be_gs call ran the line kernel model, -1: it deadlocked
extern "C" int pdeop_emu_line_last() { return g_line_last; }

struct LineHostIO {
    std::vector<std::pair<double*, const double*>> pending;
    double ldcg(const double* p) { return *p; }
    void prefetch(const void*) {}
    double ldstream(const double* p) { return *p; }
    double ldown(const double* p) { return *p; }
    void cp8(double* dst, const double* src) { pending.emplace_back(dst, src); }
    void land(size_t k) {
        *pending[k].first = *pending[k].second;
        pending[k] = pending.back();
        pending.pop_back();
    }
    void land_all() {
        for (auto& pr : pending) *pr.first = *pr.second;
        pending.clear();
    }
};

template <int D, int PS>
static int gs_line_model(const LevelDev& L, const LineGeom& g, const double* Ti, const LineStreams& S, unsigned ioff,
                         double* x, std::mt19937& rng) {
    const int C = g.C, NT = g.threads;
    struct Th {
        LineCtx<D> c;
        LineHostIO io;
        int pc = 0;   // next phase: 2n = A(n), 2n+1 = B(n)
    };
    std::vector<std::vector<double>> smem(C, std::vector<double>(g.o_end, 0.0));
    std::vector<std::vector<Th>> th(C, std::vector<Th>(NT));
    std::vector<std::vector<int>> fin_count(C, std::vector<int>(g.NS + 1, 0));   // threads that finished A(n)
    std::vector<long long> flag(C, 0);
    for (int q = 0; q < C; ++q) {
        line_smem_fill<D, PS>(L, g, Ti, q, smem[q].data(), 0, 1);
        for (int t = 0; t < NT; ++t) line_init<D>(L, g, q, t, th[q][t].c);
    }
    long long remaining = (long long)C * NT * 2 * g.NS;
    std::vector<int> order((size_t)C * NT);
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    // may thread (q, t) run its next phase now?  (thread 0's publish is part of reaching B(n))
    auto runnable = [&](int q, int t) -> bool {
        Th& me = th[q][t];
        if (me.pc >= 2 * g.NS) return false;
        const int n = me.pc >> 1;
        if ((me.pc & 1) == 0) return n == 0 || fin_count[q][n - 1] == NT;
        if (C > 1 && t == 0) {
            flag[q] = n;
            if (q > 0 && flag[q - 1] < (long long)n + 3 - kLineDelta) return false;
            if (q + 1 < C && flag[q + 1] < (long long)n + 8 + kLineDelta - g.N2) return false;
        }
        return true;
    };
    while (remaining > 0) {
        std::shuffle(order.begin(), order.end(), rng);
        bool progressed = false;
        for (int id : order) {
            const int q = id / NT, t = id % NT;
            Th& me = th[q][t];
            if (me.pc >= 2 * g.NS) continue;
            if (rng() % 4 == 0) continue;   // random stalls: threads drift apart as far as the barriers allow
            if (!runnable(q, t)) continue;
            const int n = me.pc >> 1;
            double* sm = smem[q].data();
            while (!me.io.pending.empty() && rng() % 3 == 0) me.io.land(rng() % me.io.pending.size());
            if ((me.pc & 1) == 0) {   // A(n), after WAIT(n-1)
                line_fin<D, PS>(L, g, sm, x, me.c);
                me.io.land_all();     // cp.async.wait_group 0 before ARRIVE(n)
                fin_count[q][n] += 1;
            } else {                  // B(n)
                line_pre<D, PS>(L, g, sm, S, ioff, me.c, me.io);
                line_advance<D>(g, me.c);
            }
            me.pc += 1;
            remaining -= 1;
            progressed = true;
            if (me.pc == 2 * g.NS && t == 0) flag[q] = g.NS;
        }
        if (!progressed) {   // everything runnable was stalled at random -- or nothing is runnable: a deadlock
            bool any = false;
            for (int id : order) any = any || runnable(id / NT, id % NT);
            if (!any) return -1;
        }
    }
    return 0;
}

template <int D>
static int gs_line_t(const LevelDev& L, int B, const double* T, const double* coef, const double* dinv, const double* b,
                     double* x, int nsweeps) {
    int maxn = std::max(L.N[0], std::max(L.N[1], L.N[2])) + 2 * kTabPad;
    const int ps = maxn <= 40 ? 40 : maxn <= 72 ? 72 : maxn <= 136 ? 136 : maxn <= 264 ? 264 : 0;
    LineGeom g;
    if (!line_geom(L, nsweeps, ps, 8, (size_t)227 * 1024, g_line_maxthreads > 0 ? g_line_maxthreads : kLineMaxThreads, g))
        return 0;
    std::mt19937 rng(g_line_seed);
    LineStreams S;
    line_streams(L, coef, dinv, b, x, S);
    for (int ib = 0; ib < B; ++ib) {
        const double* Ti = T + ib * tstride(L);
        const size_t o = ib * vstride(L);
        int rc;
        if (ps == 40) rc = gs_line_model<D, 40>(L, g, Ti, S, (unsigned)o, x + o, rng);
        else if (ps == 72) rc = gs_line_model<D, 72>(L, g, Ti, S, (unsigned)o, x + o, rng);
        else if (ps == 136) rc = gs_line_model<D, 136>(L, g, Ti, S, (unsigned)o, x + o, rng);
        else rc = gs_line_model<D, 264>(L, g, Ti, S, (unsigned)o, x + o, rng);
        if (rc) return -1;
    }
    return 1;
}

void be_gs(stream_t, const LevelDev& L, int B, const double* T, const double* coef, const double* dinv,
           const double* b, double* x, double*, size_t, int nsweeps, const int* done, int variant, int gs_pipe) {
    if (done && *done) return;
    if (nsweeps <= 0) return;
    g_line_last = 0;
    if (variant == 0 && gs_pipe == 5 && L.D >= 2) {
        g_line_last = L.D == 2 ? gs_line_t<2>(L, B, T, coef, dinv, b, x, nsweeps)
                               : gs_line_t<3>(L, B, T, coef, dinv, b, x, nsweeps);
        if (g_line_last != 0) return;
    }
    if (L.D == 1) gs_t<1>(L, B, T, coef, dinv, b, x, nsweeps);
    else if (L.D == 2) gs_t<2>(L, B, T, coef, dinv, b, x, nsweeps);
    else gs_t<3>(L, B, T, coef, dinv, b, x, nsweeps);
}

void be_dinv(stream_t, const LevelDev& L, int B, const double* T, const double* coef, double* dinv) {
    for (int ib = 0; ib < B; ++ib)
        for (int w = 0; w < L.G; ++w) {
            if (L.D == 1) dinv_elem<1>(L, T + ib * tstride(L), coef + ib * vstride(L), dinv + ib * vstride(L), w);
            else if (L.D == 2) dinv_elem<2>(L, T + ib * tstride(L), coef + ib * vstride(L), dinv + ib * vstride(L), w);
            else dinv_elem<3>(L, T + ib * tstride(L), coef + ib * vstride(L), dinv + ib * vstride(L), w);
        }
}

void be_zero_dense(stream_t, int B, int n, int, double* Kd) { memset(Kd, 0, (size_t)B * n * n * sizeof(double)); }

void be_dense(stream_t, const LevelDev& L, int B, const double* T, const double* coef, double* Kd) {
    const size_t n = vstride(L);
    for (int ib = 0; ib < B; ++ib)
        for (int w = 0; w < L.G; ++w) {
            if (L.D == 1) dense_elem<1>(L, T + ib * tstride(L), coef + ib * n, Kd + ib * n * n, w);
            else if (L.D == 2) dense_elem<2>(L, T + ib * tstride(L), coef + ib * n, Kd + ib * n * n, w);
            else dense_elem<3>(L, T + ib * tstride(L), coef + ib * n, Kd + ib * n * n, w);
        }
}

void* be_aux_create() { return nullptr; }
void be_aux_destroy(void*) {}
void be_cholesky(stream_t, int B, int n, int, double* Kd, double*, FgmresState* state, bool, void*) {
    for (int ib = 0; ib < B; ++ib) {
        double* A = Kd + (size_t)ib * n * n;
        for (int j = 0; j < n; ++j) {
            double d = A[(size_t)j * n + j];
            for (int k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
            if (!(d > 0.0)) {
                if (state->chol_info == 0) state->chol_info = j + 1;
                d = 1.0;
            }
            d = sqrt(d);
            A[(size_t)j * n + j] = d;
            for (int i = j + 1; i < n; ++i) {
                double v = A[(size_t)i * n + j];
                for (int k = 0; k < j; ++k) v -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
                A[(size_t)i * n + j] = v / d;
            }
        }
    }
}

void be_chol_solve(stream_t, const LevelDev& L, int B, const double* Lf, const double*, const double* rhs, double* out,
                   double* work, const int* done, bool) {
    if (done && *done) return;
    const int n = L.M * L.G;
    for (int ib = 0; ib < B; ++ib) {
        const double* A = Lf + (size_t)ib * n * n;
        double* y = work + (size_t)ib * n;
        for (int w = 0; w < L.G; ++w) to_band_elem(L, rhs + (size_t)ib * n, y, w);
        for (int i = 0; i < n; ++i) {
            double v = y[i];
            for (int k = 0; k < i; ++k) v -= A[(size_t)i * n + k] * y[k];
            y[i] = v / A[(size_t)i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double v = y[i];
            for (int k = i + 1; k < n; ++k) v -= A[(size_t)k * n + i] * y[k];
            y[i] = v / A[(size_t)i * n + i];
        }
        for (int w = 0; w < L.G; ++w) from_band_elem(L, y, out + (size_t)ib * n, w);
    }
}

void be_grads(stream_t, const LevelDev& L, int B, const double* coef, const double* rhs_nat, const double* cv,
              const double* fv, const double* bv, const double* x, const double* dz, double* d_coeffs, double* d_rhs,
              double* d_iv, double* d_cv, double* d_fv, double* d_bv) {
    const size_t n = vstride(L);
    for (int ib = 0; ib < B; ++ib) {
        const size_t oc = (size_t)ib * L.Ntot * 12, of = (size_t)ib * L.Ftot * 4;
        for (int w = 0; w < L.G; ++w) {
#define GRAD_CALL(DD)                                                                                              \
    grad_elem<DD>(L, coef + ib * n, rhs_nat + (size_t)ib * L.G, cv + oc, fv + of, bv + of, x + ib * n, dz + ib * n, \
                  d_coeffs + ib * n, d_rhs + (size_t)ib * L.G, d_cv + oc, d_fv + of, d_bv + of, w)
            if (L.D == 1) GRAD_CALL(1);
            else if (L.D == 2) GRAD_CALL(2);
            else GRAD_CALL(3);
#undef GRAD_CALL
        }
        for (int k = 0; k < L.n_init; ++k) grad_init_elem(L, dz + ib * n, d_iv + (size_t)ib * L.n_init, k);
    }
}

// ---- FGMRES vector steps --------------------------------------------------------------------------
static double nrm2(size_t n, const double* v) {
    double s = 0.0;
    for (size_t i = 0; i < n; ++i) s += v[i] * v[i];
    return sqrt(s);
}

void be_state_reset(stream_t, FgmresState* s) { memset(s, 0, sizeof(*s)); }

void be_fg_begin(stream_t, size_t n, const double* b, double* x, FgmresState* s) {
    memset(x, 0, n * sizeof(double));
    s->bnorm = nrm2(n, b);
    s->iters = 0;
    s->rnorm = 0.0;
    s->done = (s->bnorm == 0.0) ? 1 : 0;  // fgmres.py:76-78
}

void be_fg_resnorm(stream_t, size_t n, const double* r, FgmresState* s, int maxiter, double atol) {
    if (s->done) return;
    s->rnorm = nrm2(n, r);
    if (s->rnorm <= atol || s->iters >= maxiter) s->done = 1;  // fgmres.py:134
    else s->e[0] = s->rnorm;
}

void be_fg_first(stream_t, size_t n, const double* r, double* V0, FgmresState* s) {
    if (s->done) return;
    for (size_t i = 0; i < n; ++i) V0[i] = r[i] / s->rnorm;
}

void be_fg_cgs(stream_t, size_t n, int j, int restart, double* V, double* w, FgmresState* s) {
    if (s->done) return;
    double h[kMaxRestart];
    for (int k = 0; k <= j; ++k) {
        double a = 0.0;
        const double* vk = V + (size_t)k * n;
        for (size_t i = 0; i < n; ++i) a += vk[i] * w[i];
        h[k] = a;
        s->H[k * restart + j] = a;
    }
    for (int k = 0; k <= j; ++k) {
        const double* vk = V + (size_t)k * n;
        for (size_t i = 0; i < n; ++i) w[i] -= h[k] * vk[i];
    }
    const double nn = nrm2(n, w);
    s->H[(j + 1) * restart + j] = nn;
    if (j + 1 < restart) {
        double* vn = V + (size_t)(j + 1) * n;
        for (size_t i = 0; i < n; ++i) vn[i] = w[i] / nn;
    }
}

}  // namespace pdeop

#include "../../mech_nn_discovery_pde_b200/csrc/pdeop_lstsq.h"

namespace pdeop {

void be_fg_update(stream_t, size_t n, int restart, const double* Z, double* x, FgmresState* s) {
    if (s->done) return;
    hessenberg_lstsq(s->H, s->e, restart, s->y);
    for (int k = 0; k < restart; ++k) {
        const double* zk = Z + (size_t)k * n;
        const double yk = s->y[k];
        for (size_t i = 0; i < n; ++i) x[i] += yk * zk[i];
    }
    s->iters += restart;
}

void be_fg_info(stream_t, const FgmresState* s, double* info4) {
    info4[0] = s->iters;
    info4[1] = s->rnorm;
    info4[2] = s->bnorm;
    info4[3] = s->chol_info;
}

void be_fg_hess(stream_t, const FgmresState* s, int restart, double* hess_out) {
    memcpy(hess_out, s->H, sizeof(double) * (restart + 1) * restart);
}

}  // namespace pdeop
