// pdeop -- host-side orchestration: plan (closed-form index tables), multigrid V-cycle, FGMRES,
// layer forward/backward, and the C ABI of include/pdeop.h.  All device work goes through the
// be_* backend calls (pdeop_backend.h); nothing here synchronises the host with the device.
//
// Algorithm restated from the reference (paths relative to the reference repo):
//   level hierarchy             solver/multigrid.py:88-112
//   coarse operators            solver/multigrid.py:115-163, 243-285 (rediscretisation on interpolated coeffs)
//   V(nu1,nu2) cycle            solver/multigrid.py:453-498
//   restarted FGMRES            solver/fgmres.py:21-182
//   forward / backward          solver/qp_dual_sparse_multigrid_normal_kkt.py:25-162, qp_dual_dense_normal_kkt.py:23-118
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/pdeop.h"
#include "pdeop_backend.h"

using namespace pdeop;

static thread_local std::string g_err;

// ---- optional per-category timing with events on the launching stream (bench.py roofline numbers) ----
enum ProfCat {
    PC_GS_FINE = 0, PC_GS_COARSE, PC_APPLY_FINE, PC_APPLY_COARSE, PC_TRANSFER, PC_COARSE_SOLVE, PC_FACTOR,
    PC_KRYLOV, PC_SETUP, PC_GRADS, PC_LAYOUT, PC_COUNT
};
struct ProfRec { int cat; void* a; void* b; };
// per-plan instrumentation state (a plan is driven by one host thread at a time; two plans never share state)
struct ProfState {
    bool on = false;
    std::vector<ProfRec> recs;
    std::vector<void*> pool;
    void* event() {
        if (!pool.empty()) { void* e = pool.back(); pool.pop_back(); return e; }
        return be_event_create();
    }
    ~ProfState() {
        for (auto& r : recs) { be_event_destroy(r.a); be_event_destroy(r.b); }
        for (void* e : pool) be_event_destroy(e);
    }
};
struct ProfScope {
    ProfState* ps; int cat; void* a = nullptr; stream_t st;
    ProfScope(ProfState& p, int c, stream_t s) : ps(&p), cat(c), st(s) {
        if (ps->on) { a = ps->event(); be_event_record(a, st); }
    }
    ~ProfScope() {
        if (a) { void* b = ps->event(); be_event_record(b, st); ps->recs.push_back({cat, a, b}); }
    }
};
extern "C" long long pdeop_launch_count(void) { return be_launch_count(); }

static int fail(const std::string& msg) {
    g_err = msg;
    return 1;
}

static int check_backend() {
    char buf[512];
    if (be_last_error(buf, sizeof(buf))) return fail(std::string("backend error: ") + buf);
    return 0;
}

struct LevelHost {
    LevelDev dev;
    std::vector<void*> owned;  // device allocations
    size_t off_T = 0, off_coef = 0, off_dinv = 0;   // persist offsets (doubles)
    size_t off_x = 0, off_b = 0, off_r = 0;  // scratch offsets (doubles), levels >= 1
};

struct pdeop_plan {
    int D = 0, M = 0, B = 0, n_grid = 0, dsf = 0, order = 2;
    std::vector<LevelHost> lev;
    size_t persist_doubles = 0;
    size_t off_Kd = 0;  // persist offset of the coarsest dense matrix / factor
    size_t off_Linv = 0;  // persist offset of the inverse diagonal blocks of the factor
    int nc = 0;         // coarsest unknowns per instance
    // The dense nc x nc array lives in the (per-layer, reused) scratch buffer instead of the (per-forward-call)
    // persist buffer when the chain solver is active: it is then dead after operator set-up.
    bool kd_in_scratch = false;
    // Kernel-variant choices, fixed when the plan is created (the buffer layout depends on use_chain, so the
    // layout and the solver choice come from the same snapshot); gs_pipe may be changed per plan for A/B tests.
    bool use_chain = true;
    int gs_pipe = 2;
    int device = -1;   // CUDA device the plan's tables live on; every entry point checks it is current
    void* aux = nullptr;   // backend helper objects owned by the plan (side stream + events of the factorisation)
    ProfState prof;
};

template <class T>
static T* upload_vec(LevelHost& lh, const std::vector<T>& v) {
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    void* p = be_alloc(bytes);
    if (!v.empty()) be_upload(p, v.data(), v.size() * sizeof(T));
    lh.owned.push_back(p);
    return (T*)p;
}

static int build_level(pdeop_plan* pl, LevelHost& lh, const int* dims, int n_iv, const int* iv_desc) {
    const int D = pl->D;
    LevelDev& L = lh.dev;
    memset(&L, 0, sizeof(L));
    for (int a = 0; a < 3; ++a) L.N[a] = 1;
    for (int c = 0; c < D; ++c) L.N[3 - D + c] = dims[c];
    for (int a = 0; a < 3; ++a)
        if (L.N[a] > 1023) return fail("grid extent above 1023 not supported");
    L.D = D;
    L.order = pl->order;
    L.M = 1 + 2 * D;
    L.G = L.N[0] * L.N[1] * L.N[2];
    L.S = L.N[0] + L.N[1] + L.N[2] - 2;
    L.P = kTabPitch;
    if ((long long)L.M * L.G >= (1LL << 31)) return fail("grid too large for 32-bit vector indexing");
    L.Ntot = 0;
    L.Ftot = 0;
    for (int c = 0; c < D; ++c) {
        if (dims[c] < 6) return fail("every grid extent must be at least 6 (one-sided 5-point stencils)");
        L.cvoff[c] = L.Ntot;
        L.fvoff[c] = L.Ftot;
        L.Ntot += dims[c];
        L.Ftot += dims[c] - 1;
    }
    const int N0 = L.N[0], N1 = L.N[1], N2 = L.N[2];
    std::vector<int> coord(L.G), flags(L.G, 0), hstart(L.S + 1), rowbase((size_t)(L.S + 8) * N0 + 8, 0);
    std::vector<int> nat2wave(L.G);
    int w = 0;
    for (int s = 0; s < L.S; ++s) {
        hstart[s] = w;
        for (int i0 = 0; i0 < N0; ++i0) {
            int lo = std::max(0, s - i0 - (N2 - 1));
            int hi = std::min(N1 - 1, s - i0);
            rowbase[4 + (size_t)(s + 4) * N0 + i0] = w - lo;
            for (int i1 = lo; i1 <= hi; ++i1) {
                int i2 = s - i0 - i1;
                coord[w] = i0 | (i1 << 10) | (i2 << 20);
                nat2wave[(i0 * N1 + i1) * N2 + i2] = w;
                ++w;
            }
        }
    }
    hstart[L.S] = w;
    if (w != L.G) return fail("internal: wave enumeration mismatch");
    // equation rows: skip first-axis index 0 and both ends of every other axis (lp_pde_central_diff.py:228-235)
    int n_eq = 0;
    for (int ww = 0; ww < L.G; ++ww) {
        int idx[3];
        unpack_coord(coord[ww], idx[0], idx[1], idx[2]);
        bool eq = idx[3 - D] != 0;
        for (int c = 1; c < D; ++c) {
            int i = idx[3 - D + c];
            if (i == 0 || i == dims[c] - 1) eq = false;
        }
        if (eq) {
            flags[ww] |= 1;
            coord[ww] |= 1 << 30;
            ++n_eq;
        }
    }
    L.n_eq = n_eq;
    // initial / boundary rows in construction order: per spec, grid C order inside the box (:1008-1033)
    std::vector<int> init_w, init_m;
    for (int k = 0; k < n_iv; ++k) {
        const int* dsc = iv_desc + (size_t)k * (1 + 2 * D);
        int mi = dsc[0];
        if (mi < 0 || mi >= 1 + pl->order * D) return fail("initial-condition multi-index out of range");
        int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        for (int c = 0; c < D; ++c) {
            lo[3 - D + c] = std::max(0, dsc[1 + c]);
            hi[3 - D + c] = std::min(dims[c] - 1, dsc[1 + D + c]);
        }
        for (int i0 = lo[0]; i0 <= hi[0]; ++i0)
            for (int i1 = lo[1]; i1 <= hi[1]; ++i1)
                for (int i2 = lo[2]; i2 <= hi[2]; ++i2) {
                    int ww = nat2wave[(i0 * N1 + i1) * N2 + i2];
                    int cnt = (flags[ww] >> (4 + 2 * mi)) & 3;
                    if (cnt == 3) return fail("more than 3 initial rows on one variable");
                    flags[ww] += 1 << (4 + 2 * mi);
                    init_w.push_back(ww);
                    init_m.push_back(mi);
                }
    }
    L.n_init = (int)init_w.size();
    // band ordering for the dense path: longest axis outermost
    {
        int outer = 0;
        for (int a = 1; a < 3; ++a)
            if (L.N[a] > L.N[outer]) outer = a;
        const int in1 = outer == 0 ? 1 : 0, in2 = outer == 2 ? 1 : 2;
        const int inner = L.N[in1] * L.N[in2];
        std::vector<int> band(L.G);
        for (int ww = 0; ww < L.G; ++ww) {
            int idx[3];
            unpack_coord(coord[ww], idx[0], idx[1], idx[2]);
            band[ww] = idx[outer] * inner + idx[in1] * L.N[in2] + idx[in2];
        }
        L.band = upload_vec(lh, band);
        L.bw = 4 * inner * L.M + L.M - 1;
        if (L.bw > L.M * L.G - 1) L.bw = L.M * L.G - 1;
    }
    L.coord = upload_vec(lh, coord);
    L.flags = upload_vec(lh, flags);
    L.hstart = upload_vec(lh, hstart);
    L.rowbase = upload_vec(lh, rowbase) + 4;   // 4 spare ints either side: branch-free neighbour lookups
    L.init_w = upload_vec(lh, init_w);
    L.init_m = upload_vec(lh, init_m);
    return 0;
}

extern "C" int pdeop_plan_create(int d, const int* dims, int order, int batch, int n_grid, int downsample_first,
                                 int n_iv, const int* iv_desc, pdeop_plan** out) {
    return pdeop_plan_create_ex(d, dims, order, batch, n_grid, downsample_first, n_iv, iv_desc, nullptr, out);
}

extern "C" int pdeop_plan_create_ex(int d, const int* dims, int order, int batch, int n_grid, int downsample_first,
                                    int n_iv, const int* iv_desc, const pdeop_plan_opts* opts, pdeop_plan** out) {
    if (!out) return fail("null out");
    *out = nullptr;
    if (d < 1 || d > 3) return fail("dimension must be 1, 2 or 3");
    if (order != 1 && order != 2) return fail("total order must be 1 or 2");
    if (batch < 1 || n_grid < 1) return fail("batch and n_grid must be positive");
    if (batch > 65535) return fail("batch above 65535 not supported (instances map to gridDim.y)");
    if (opts && opts->evolution) return fail("evolution=True equation rows are not implemented");
    pdeop_plan* pl = new pdeop_plan();
    pl->device = be_current_device();
    pl->gs_pipe = (opts && opts->gs_pipe >= 0) ? opts->gs_pipe : be_default_gs_pipe();
    if (pl->gs_pipe > 5) { delete pl; return fail("gs_pipe must be 0..5"); }
    const bool want_chain = (opts && opts->chain >= 0) ? opts->chain != 0 : be_default_chain();
    pl->D = d;
    pl->M = 1 + 2 * d;
    pl->B = batch;
    pl->n_grid = n_grid;
    pl->dsf = downsample_first;
    pl->order = order;
    pl->lev.resize(n_grid);
    int cur[3];
    for (int c = 0; c < d; ++c) cur[c] = dims[c];
    size_t poff = 0;
    for (int l = 0; l < n_grid; ++l) {
        LevelHost& lh = pl->lev[l];
        if (build_level(pl, lh, cur, n_iv, iv_desc + (size_t)l * n_iv * (1 + 2 * d))) {
            pdeop_plan_destroy(pl);
            return 1;
        }
        const LevelDev& L = lh.dev;
        lh.off_T = poff;
        poff += (size_t)batch * d * kTabEntries * L.P;
        lh.off_coef = poff;
        poff += (size_t)batch * L.M * L.G;
        lh.off_dinv = poff;
        poff += (size_t)batch * L.M * L.G;
        for (int c = 0; c < d; ++c)  // multigrid.py:99-102
            if (c > 0 || downsample_first) cur[c] /= 2;
    }
    const LevelDev& Lc = pl->lev[n_grid - 1].dev;
    pl->nc = Lc.M * Lc.G;
    pl->off_Kd = poff;
    pl->use_chain = be_chain_layout(pl->nc, Lc.bw).use && want_chain;
    pl->kd_in_scratch = pl->use_chain;
    if (!pl->kd_in_scratch) poff += (size_t)batch * pl->nc * pl->nc;
    pl->off_Linv = poff;
    poff += be_chol_linv_doubles(batch, pl->nc, Lc.bw, pl->use_chain);   // inverse diagonal blocks (+ transposes, + scaled band)
    pl->persist_doubles = poff;
    pl->aux = be_aux_create();
    if (check_backend()) {   // a failed table allocation or upload
        pdeop_plan_destroy(pl);
        return 1;
    }
    *out = pl;
    return 0;
}

extern "C" void pdeop_plan_destroy(pdeop_plan* pl) {
    if (!pl) return;
    for (auto& lh : pl->lev)
        for (void* p : lh.owned) be_free(p);
    if (pl->aux) be_aux_destroy(pl->aux);
    delete pl;
}

// scratch layout (doubles): [state | atb | x | w | V[restart] | Z[restart] | per level>=1: x,b,r | cwork]
struct Scratch {
    FgmresState* state;
    double *atb, *x, *w, *V, *Z, *cwork, *gs_stash, *Kd;
    size_t gs_stash_stride;
    std::vector<double*> lx, lb, lr;
    size_t n0;
    // converged mode: per-instance scalars (8 arrays of B) and lambda_max estimates (one array of B per level)
    double* bscal;
    double* lam;
};

static size_t state_doubles() { return kStateAreaDoubles; }

static size_t scratch_doubles(const pdeop_plan* pl, int restart) {
    const LevelDev& L0 = pl->lev[0].dev;
    size_t n0 = (size_t)pl->B * L0.M * L0.G;
    size_t tot = state_doubles() + n0 * (3 + 2 * (size_t)std::max(restart, 1));
    if (pl->kd_in_scratch) tot += (size_t)pl->B * pl->nc * pl->nc;
    for (int l = 1; l < pl->n_grid; ++l) tot += 3 * (size_t)pl->B * pl->lev[l].dev.M * pl->lev[l].dev.G;
    tot += 2 * (size_t)pl->B * pl->nc;
    tot += (size_t)pl->B * be_gs_stash_doubles(L0);
    tot += (size_t)pl->B * (8 + pl->n_grid);
    return tot;
}

static Scratch carve(const pdeop_plan* pl, void* scratch, int restart) {
    Scratch s;
    double* p = (double*)scratch;
    s.state = (FgmresState*)p;
    p += state_doubles();
    s.Kd = nullptr;
    if (pl->kd_in_scratch) {   // first, so that its place does not depend on `restart`
        s.Kd = p;
        p += (size_t)pl->B * pl->nc * pl->nc;
    }
    const LevelDev& L0 = pl->lev[0].dev;
    s.n0 = (size_t)pl->B * L0.M * L0.G;
    s.atb = p; p += s.n0;
    s.x = p; p += s.n0;
    s.w = p; p += s.n0;
    s.V = p; p += s.n0 * std::max(restart, 1);
    s.Z = p; p += s.n0 * std::max(restart, 1);
    s.lx.assign(pl->n_grid, nullptr);
    s.lb.assign(pl->n_grid, nullptr);
    s.lr.assign(pl->n_grid, nullptr);
    for (int l = 1; l < pl->n_grid; ++l) {
        size_t nl = (size_t)pl->B * pl->lev[l].dev.M * pl->lev[l].dev.G;
        s.lx[l] = p; p += nl;
        s.lb[l] = p; p += nl;
        s.lr[l] = p; p += nl;
    }
    s.cwork = p; p += 2 * (size_t)pl->B * pl->nc;
    s.gs_stash = p;
    s.gs_stash_stride = be_gs_stash_doubles(L0);
    p += (size_t)pl->B * s.gs_stash_stride;
    s.bscal = p; p += (size_t)pl->B * 8;
    s.lam = p;
    return s;
}

extern "C" int pdeop_plan_query(const pdeop_plan* pl, int what, int level, long long* out) {
    if (!pl || !out) return fail("null argument");
    if (what >= PDEOP_Q_DIM0 && what < PDEOP_Q_DIM0 + 3) {
        if (level < 0 || level >= pl->n_grid) return fail("level out of range");
        int c = what - PDEOP_Q_DIM0;
        if (c >= pl->D) return fail("axis out of range");
        *out = pl->lev[level].dev.N[3 - pl->D + c];
        return 0;
    }
    if (what == PDEOP_Q_NLEVELS) { *out = pl->n_grid; return 0; }
    if (what == PDEOP_Q_PERSIST_BYTES) { *out = (long long)(pl->persist_doubles * sizeof(double)); return 0; }
    if (what == PDEOP_Q_SCRATCH_BYTES) { *out = (long long)(scratch_doubles(pl, level) * sizeof(double)); return 0; }
    if (level < 0 || level >= pl->n_grid) return fail("level out of range");
    const LevelDev& L = pl->lev[level].dev;
    switch (what) {
        case PDEOP_Q_G: *out = L.G; return 0;
        case PDEOP_Q_M: *out = L.M; return 0;
        case PDEOP_Q_N_EQ: *out = L.n_eq; return 0;
        case PDEOP_Q_N_INIT: *out = L.n_init; return 0;
        case PDEOP_Q_NTOT: *out = L.Ntot; return 0;
        case PDEOP_Q_FTOT: *out = L.Ftot; return 0;
    }
    return fail("unknown query");
}

// every compute entry point: the plan's tables live on one device, which must be the calling thread's current one
static int check_plan(const pdeop_plan* pl) {
    if (!pl) return fail("null plan");
    const int cur = be_current_device();
    if (cur != pl->device) {
        char buf[160];
        snprintf(buf, sizeof(buf), "plan was created on device %d but device %d is current; cudaSetDevice (torch.cuda."
                 "device) to the plan's device before calling", pl->device, cur);
        return fail(buf);
    }
    return 0;
}

extern "C" int pdeop_plan_set_tuning(pdeop_plan* pl, int key, int value) {
    if (!pl) return fail("null plan");
    if (key == 0) {
        if (value < 0 || value > 5) return fail("gs_pipe must be 0..5");
        pl->gs_pipe = value;
        return 0;
    }
    // key 1 ("chain") is a property of the plan's buffer layout: it can only be chosen at creation (PDEOP_CHAIN)
    if (key == 1) return fail("the chain-solver choice is fixed when the plan is created (environment PDEOP_CHAIN)");
    return fail("unknown tuning key");
}
extern "C" int pdeop_plan_get_tuning(const pdeop_plan* pl, int key, int* value) {
    if (!pl || !value) return fail("null argument");
    if (key == 0) { *value = pl->gs_pipe; return 0; }
    if (key == 1) { *value = pl->use_chain ? 1 : 0; return 0; }
    return fail("unknown tuning key");
}
extern "C" int pdeop_plan_profile_enable(pdeop_plan* pl, int on) {
    if (!pl) return fail("null plan");
    ProfState& p = pl->prof;
    p.on = on != 0;
    for (auto& r : p.recs) { p.pool.push_back(r.a); p.pool.push_back(r.b); }
    p.recs.clear();
    return 0;
}
extern "C" int pdeop_plan_profile_collect(pdeop_plan* pl, double* ms, long long* counts, int ncat) {
    if (!pl) return -1;
    ProfState& p = pl->prof;
    for (int i = 0; i < ncat; ++i) { ms[i] = 0.0; counts[i] = 0; }
    for (auto& r : p.recs) {
        if (r.cat < ncat) { ms[r.cat] += be_event_elapsed_ms(r.a, r.b); counts[r.cat] += 1; }
        p.pool.push_back(r.a);
        p.pool.push_back(r.b);
    }
    p.recs.clear();
    return PC_COUNT;
}

extern "C" const char* pdeop_last_error(void) { return g_err.c_str(); }
extern "C" const char* pdeop_backend_name(void) { return be_name(); }

static double* P_T(const pdeop_plan* pl, void* persist, int l) { return (double*)persist + pl->lev[l].off_T; }
static double* P_coef(const pdeop_plan* pl, void* persist, int l) { return (double*)persist + pl->lev[l].off_coef; }
static double* P_dinv(const pdeop_plan* pl, void* persist, int l) { return (double*)persist + pl->lev[l].off_dinv; }
static double* P_Kd(const pdeop_plan* pl, void* persist, const Scratch& sc) {
    return pl->kd_in_scratch ? sc.Kd : (double*)persist + pl->off_Kd;
}
static double* P_Linv(const pdeop_plan* pl, void* persist) { return (double*)persist + pl->off_Linv; }

// Operator set-up: level-0 coefficients into wave layout, coarse coefficients by linear interpolation
// of the finer level's (multigrid.py:243-256), axis tables of every level from that level's line
// values, dense coarsest K and its Cholesky factor (multigrid.py:216-218, 438-440).
static void setup_operator(pdeop_plan* pl, const double* coeffs, const double* const* cv, const double* const* fv,
                           const double* const* bv, void* persist, Scratch& sc, stream_t st) {
    const int B = pl->B;
    be_state_reset(st, sc.state);
    {
        ProfScope ps(pl->prof, PC_SETUP, st);
        be_pack(st, pl->lev[0].dev, B, coeffs, P_coef(pl, persist, 0));
        for (int l = 0; l < pl->n_grid; ++l) {
            const LevelDev& L = pl->lev[l].dev;
            if (l > 0)
                be_interp(st, pl->lev[l - 1].dev, L, B, L.M, P_coef(pl, persist, l - 1), P_coef(pl, persist, l), 0,
                          nullptr);
            be_build_tables(st, L, B, cv[l], fv[l], bv[l], P_T(pl, persist, l));
            be_dinv(st, L, B, P_T(pl, persist, l), P_coef(pl, persist, l), P_dinv(pl, persist, l));
        }
        const int lc = pl->n_grid - 1;
        double* Kd = P_Kd(pl, persist, sc);
        be_zero_dense(st, B, pl->nc, pl->lev[lc].dev.bw, Kd);
        be_dense(st, pl->lev[lc].dev, B, P_T(pl, persist, lc), P_coef(pl, persist, lc), Kd);
    }
    ProfScope pf(pl->prof, PC_FACTOR, st);
    be_cholesky(st, B, pl->nc, pl->lev[pl->n_grid - 1].dev.bw, P_Kd(pl, persist, sc), P_Linv(pl, persist), sc.state, pl->use_chain,
                pl->aux);
}

static void vcycle(pdeop_plan* pl, const pdeop_solver_cfg* cfg, void* persist, Scratch& sc, int l, const double* b,
                   double* x, double* rtmp, stream_t st) {
    const int B = pl->B;
    const LevelDev& L = pl->lev[l].dev;
    const int* done = &sc.state->done;
    {
        ProfScope ps(pl->prof, l == 0 ? PC_GS_FINE : PC_GS_COARSE, st);
        be_gs(st, L, B, P_T(pl, persist, l), P_coef(pl, persist, l), P_dinv(pl, persist, l), b, x, sc.gs_stash, sc.gs_stash_stride,
              cfg->gs_pre, done, cfg->gs_variant, pl->gs_pipe);
    }
    {
        ProfScope ps(pl->prof, l == 0 ? PC_APPLY_FINE : PC_APPLY_COARSE, st);
        be_apply_k(st, L, B, P_T(pl, persist, l), P_coef(pl, persist, l), x, b, rtmp, 1, done);
    }
    const LevelDev& Lc = pl->lev[l + 1].dev;
    {
        ProfScope ps(pl->prof, PC_TRANSFER, st);
        be_interp(st, L, Lc, B, L.M, rtmp, sc.lb[l + 1], 0, done);
    }
    if (l + 1 == pl->n_grid - 1) {
        ProfScope ps(pl->prof, PC_COARSE_SOLVE, st);
        be_chol_solve(st, Lc, B, P_Kd(pl, persist, sc), P_Linv(pl, persist), sc.lb[l + 1], sc.lx[l + 1], sc.cwork, done, pl->use_chain);
    } else {
        be_zero(st, sc.lx[l + 1], (size_t)B * Lc.M * Lc.G * sizeof(double));
        vcycle(pl, cfg, persist, sc, l + 1, sc.lb[l + 1], sc.lx[l + 1], sc.lr[l + 1], st);
    }
    {
        ProfScope ps(pl->prof, PC_TRANSFER, st);
        be_interp(st, Lc, L, B, L.M, sc.lx[l + 1], x, 1, done);
    }
    ProfScope ps(pl->prof, l == 0 ? PC_GS_FINE : PC_GS_COARSE, st);
    be_gs(st, L, B, P_T(pl, persist, l), P_coef(pl, persist, l), P_dinv(pl, persist, l), b, x, sc.gs_stash, sc.gs_stash_stride,
          cfg->gs_post, done, cfg->gs_variant, pl->gs_pipe);
}

// z = V-cycle(b) from a zero initial guess (multigrid.py:490-498)
static void vcycle_start(pdeop_plan* pl, const pdeop_solver_cfg* cfg, void* persist, Scratch& sc, const double* b,
                         double* z, double* rtmp, stream_t st) {
    be_zero(st, z, sc.n0 * sizeof(double));
    for (int k = 0; k < cfg->mg_steps; ++k) vcycle(pl, cfg, persist, sc, 0, b, z, rtmp, st);
}

// Restarted FGMRES(restart) with classical Gram-Schmidt and batch-global norms (fgmres.py:21-182).
// The convergence test runs on the device (state->done); once set every later launch is a no-op.
static void fgmres(pdeop_plan* pl, const pdeop_solver_cfg* cfg, void* persist, Scratch& sc, const double* b,
                   double* x, stream_t st) {
    const int B = pl->B;
    const LevelDev& L0 = pl->lev[0].dev;
    const double* T0 = P_T(pl, persist, 0);
    const double* c0 = P_coef(pl, persist, 0);
    const int m = cfg->restart;
    const int* done = &sc.state->done;
    {
        ProfScope ps(pl->prof, PC_KRYLOV, st);
        be_fg_begin(st, sc.n0, b, x, sc.state);
    }
    const int ncycles = (cfg->max_iter + m - 1) / m;
    for (int cyc = 0;; ++cyc) {
        {
            ProfScope ps(pl->prof, PC_APPLY_FINE, st);
            be_apply_k(st, L0, B, T0, c0, x, b, sc.w, 1, done);
        }
        {
            ProfScope ps(pl->prof, PC_KRYLOV, st);
            be_fg_resnorm(st, sc.n0, sc.w, sc.state, cfg->max_iter, cfg->atol);
            if (cyc == ncycles) break;
            be_fg_first(st, sc.n0, sc.w, sc.V, sc.state);
        }
        for (int j = 0; j < m; ++j) {
            double* zj = sc.Z + (size_t)j * sc.n0;
            vcycle_start(pl, cfg, persist, sc, sc.V + (size_t)j * sc.n0, zj, sc.w, st);
            {
                ProfScope ps(pl->prof, PC_APPLY_FINE, st);
                be_apply_k(st, L0, B, T0, c0, zj, nullptr, sc.w, 0, done);
            }
            ProfScope ps(pl->prof, PC_KRYLOV, st);
            be_fg_cgs(st, sc.n0, j, m, sc.V, sc.w, sc.state);
        }
        ProfScope ps(pl->prof, PC_KRYLOV, st);
        be_fg_update(st, sc.n0, m, sc.Z, x, sc.state);
    }
}

// =================================================================================================
// Converged mode: PCG with per-instance masks (solver/cg.py:51-147) preconditioned by a
// symmetric V-cycle with a polynomial smoother (weighted Jacobi, solver/multigrid.py:407-416, or Chebyshev)
// =================================================================================================
static int check_pcg(const pdeop_pcg_cfg* c) {
    if (!c) return fail("null cfg");
    if (c->max_iter < 0 || c->sweeps < 1 || c->sweeps > 16 || c->power_iters < 1) return fail("bad PCG knob");
    if (c->smoother != 0 && c->smoother != 1) return fail("smoother must be 0 (Jacobi) or 1 (Chebyshev)");
    if (!(c->rtol >= 0.0) || !(c->jacobi_w > 0.0)) return fail("bad PCG tolerance / weight");
    return 0;
}


// lam[l][b] ~ lambda_max(D^-1 K_l) by power iteration from a positive start vector (v = dinv .* dinv)
static void estimate_lambda(pdeop_plan* pl, const pdeop_pcg_cfg* cfg, void* persist, Scratch& sc, stream_t st) {
    const int B = pl->B;
    for (int l = 0; l + 1 < pl->n_grid; ++l) {
        const LevelDev& L = pl->lev[l].dev;
        const size_t n = (size_t)L.M * L.G;
        double* v = l == 0 ? sc.V : sc.lx[l];
        double* Kv = l == 0 ? sc.w : sc.lr[l];
        be_zero(st, v, (size_t)B * n * sizeof(double));
        be_zero(st, Kv, (size_t)B * n * sizeof(double));
        be_poly_update(st, n, B, v, Kv, P_dinv(pl, persist, l), P_dinv(pl, persist, l), 0.0, 1.0, nullptr, 0, nullptr);
        double* lam = sc.lam + (size_t)l * B;
        for (int it = 0; it < cfg->power_iters; ++it) {
            be_apply_k(st, L, B, P_T(pl, persist, l), P_coef(pl, persist, l), v, nullptr, Kv, 0, nullptr);
            be_power_step(st, n, B, v, Kv, P_dinv(pl, persist, l), lam, sc.bscal + 7 * (size_t)B);
        }
    }
}

// x <- x + p(D^-1 K)(b - K x): `sweeps` steps of the polynomial smoother on level l (r, d: temporaries).
// Chebyshev over [1.1 lambda_max / 30, 1.1 lambda_max] (coefficients in units of lambda_max, the kernel divides by
// the instance's estimate), or weighted Jacobi x += w D^-1 r with w = min(jacobi_w, 1.8 / lambda_max).
static void smooth_poly(pdeop_plan* pl, const pdeop_pcg_cfg* cfg, void* persist, Scratch& sc, int l, const double* b,
                        double* x, double* r, double* d, bool x_is_zero, stream_t st) {
    const int B = pl->B;
    const LevelDev& L = pl->lev[l].dev;
    const size_t n = (size_t)L.M * L.G;
    const int* done = &sc.state->done;
    const double* lam = sc.lam + (size_t)l * B;
    const double* dinv = P_dinv(pl, persist, l);
    const double lmax = 1.1, lmin = 1.1 / (cfg->cheb_ratio > 1.0 ? cfg->cheb_ratio : 30.0);
    const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
    double rho_prev = 1.0 / sigma;
    ProfScope ps(pl->prof, l == 0 ? PC_GS_FINE : PC_GS_COARSE, st);
    for (int k = 0; k < cfg->sweeps; ++k) {
        const double* res = r;
        if (k == 0 && x_is_zero) res = b;   // r = b - K 0
        else be_apply_k(st, L, B, P_T(pl, persist, l), P_coef(pl, persist, l), x, b, r, 1, done);
        if (cfg->smoother == 0) {
            if (k == 0) be_zero(st, d, (size_t)B * n * sizeof(double));
            be_poly_update(st, n, B, x, d, res, dinv, 0.0, cfg->jacobi_w, lam, 2, done);
        } else if (k == 0) {
            be_zero(st, d, (size_t)B * n * sizeof(double));
            be_poly_update(st, n, B, x, d, res, dinv, 0.0, 1.0 / theta, lam, 1, done);
        } else {
            const double rho = 1.0 / (2.0 * sigma - rho_prev);
            be_poly_update(st, n, B, x, d, res, dinv, rho * rho_prev, 2.0 * rho / delta, lam, 1, done);
            rho_prev = rho;
        }
    }
}

// z = M^-1 b: symmetric V-cycle from a zero guess (pre- and post-smoother are the same polynomial, R = P^T)
static void vcycle_sym(pdeop_plan* pl, const pdeop_pcg_cfg* cfg, void* persist, Scratch& sc, int l, const double* b,
                       double* x, double* r, double* d, stream_t st) {
    const int B = pl->B;
    const LevelDev& L = pl->lev[l].dev;
    const size_t n = (size_t)L.M * L.G;
    const int* done = &sc.state->done;
    be_zero(st, x, (size_t)B * n * sizeof(double));
    smooth_poly(pl, cfg, persist, sc, l, b, x, r, d, true, st);
    {
        ProfScope ps(pl->prof, l == 0 ? PC_APPLY_FINE : PC_APPLY_COARSE, st);
        be_apply_k(st, L, B, P_T(pl, persist, l), P_coef(pl, persist, l), x, b, r, 1, done);
    }
    const LevelDev& Lc = pl->lev[l + 1].dev;
    {
        ProfScope ps(pl->prof, PC_TRANSFER, st);
        be_restrict_t(st, L, Lc, B, L.M, r, sc.lb[l + 1], done);
    }
    if (l + 1 == pl->n_grid - 1) {
        ProfScope ps(pl->prof, PC_COARSE_SOLVE, st);
        be_chol_solve(st, Lc, B, P_Kd(pl, persist, sc), P_Linv(pl, persist), sc.lb[l + 1], sc.lx[l + 1], sc.cwork, done,
                      pl->use_chain);
    } else {
        // level l+1 temporaries: r = lr[l+1], d = slice l+1 of Z (the reference mode's Krylov basis, idle here)
        vcycle_sym(pl, cfg, persist, sc, l + 1, sc.lb[l + 1], sc.lx[l + 1], sc.lr[l + 1],
                   sc.Z + (size_t)(l + 1) * sc.n0, st);
    }
    {
        ProfScope ps(pl->prof, PC_TRANSFER, st);
        be_interp(st, Lc, L, B, L.M, sc.lx[l + 1], x, 1, done);
    }
    smooth_poly(pl, cfg, persist, sc, l, b, x, r, d, false, st);
}

// x = K^-1 b per instance to ||r_b|| <= rtol ||b_b||.  Vectors: b = sc.atb, x = sc.x, Ap = sc.w, r/z/p = V[0..2],
// level-0 smoother temporaries V[3] (r) and Z[0] (d).
static void pcg(pdeop_plan* pl, const pdeop_pcg_cfg* cfg, void* persist, Scratch& sc, const double* b, double* x,
                stream_t st) {
    const int B = pl->B;
    const LevelDev& L0 = pl->lev[0].dev;
    const size_t n = (size_t)L0.M * L0.G;
    const int* done = &sc.state->done;
    double *r = sc.V, *z = sc.V + sc.n0, *p = sc.V + 2 * sc.n0, *sr = sc.V + 3 * sc.n0, *sd = sc.Z;
    double *active = sc.bscal, *rz = active + B, *rz_new = rz + B, *pAp = rz_new + B, *rr = pAp + B, *bnorm = rr + B;
    estimate_lambda(pl, cfg, persist, sc, st);
    be_zero(st, x, sc.n0 * sizeof(double));
    be_zero(st, sc.bscal, sizeof(double) * 8 * B);
    be_bdot(st, n, B, b, b, rr, nullptr);
    be_pcg_scalars(st, B, active, rz, rz_new, pAp, rr, bnorm, cfg->rtol, sc.state, 1);
    be_copy(st, r, b, sc.n0 * sizeof(double));
    be_zero(st, p, sc.n0 * sizeof(double));
    for (int it = 0; it < cfg->max_iter; ++it) {
        vcycle_sym(pl, cfg, persist, sc, 0, r, z, sr, sd, st);
        {
            ProfScope pk(pl->prof, PC_KRYLOV, st);
            be_bdot(st, n, B, r, z, rz_new, done);
            be_pcg_p(st, n, B, p, z, rz_new, rz, active, done);   // first iteration: rz = 0 => beta = 0, p = z
        }
        {
            ProfScope pa(pl->prof, PC_APPLY_FINE, st);
            be_apply_k(st, L0, B, P_T(pl, persist, 0), P_coef(pl, persist, 0), p, nullptr, sc.w, 0, done);
        }
        ProfScope pk(pl->prof, PC_KRYLOV, st);
        be_bdot(st, n, B, p, sc.w, pAp, done);
        be_pcg_xr(st, n, B, x, r, p, sc.w, rz_new, pAp, active, rr, done);   // alpha = (r.z) / (p.Kp)
        be_pcg_scalars(st, B, active, rz, rz_new, pAp, rr, bnorm, cfg->rtol, sc.state, 0);
    }
}

static int check_cfg(const pdeop_solver_cfg* cfg) {
    if (!cfg) return fail("null cfg");
    if (cfg->restart < 1 || cfg->restart > kMaxRestart) return fail("restart must be in [1,32]");
    if (cfg->max_iter < 0 || cfg->gs_pre < 0 || cfg->gs_post < 0 || cfg->mg_steps < 0) return fail("negative knob");
    return 0;
}

extern "C" int pdeop_mg_setup(pdeop_plan* pl, const double* coeffs, const double* const* cv, const double* const* fv,
                              const double* const* bv, void* persist, void* scratch, double* info_out, void* stream) {
    if (check_plan(pl)) return 1;
    Scratch sc = carve(pl, scratch, 1);
    setup_operator(pl, coeffs, cv, fv, bv, persist, sc, stream);
    if (info_out) be_fg_info(stream, sc.state, info_out);
    return check_backend();
}

extern "C" int pdeop_mg_forward(pdeop_plan* pl, const pdeop_solver_cfg* cfg, const double* coeffs, const double* rhs,
                                const double* iv_rhs, const double* const* cv, const double* const* fv,
                                const double* const* bv, void* persist, void* scratch, double* x_out,
                                double* info_out, void* stream) {
    if (check_plan(pl)) return 1;
    if (pl->n_grid < 2) return fail("multigrid path needs n_grid >= 2");
    if (check_cfg(cfg)) return 1;
    Scratch sc = carve(pl, scratch, cfg->restart);
    setup_operator(pl, coeffs, cv, fv, bv, persist, sc, stream);
    be_atb(stream, pl->lev[0].dev, pl->B, P_coef(pl, persist, 0), rhs, iv_rhs, sc.atb);
    fgmres(pl, cfg, persist, sc, sc.atb, sc.x, stream);
    be_unpack(stream, pl->lev[0].dev, pl->B, sc.x, x_out);
    if (info_out) be_fg_info(stream, sc.state, info_out);
    return check_backend();
}

static void run_grads(pdeop_plan* pl, void* persist, Scratch& sc, const double* rhs, const double* cv0,
                      const double* fv0, const double* bv0, const double* x_api, const double* dz_wave,
                      double* d_coeffs, double* d_rhs, double* d_iv, double* d_cv, double* d_fv, double* d_bv,
                      stream_t st) {
    const LevelDev& L0 = pl->lev[0].dev;
    const int B = pl->B;
    double* xw = sc.V;  // Krylov basis no longer needed
    ProfScope ps(pl->prof, PC_GRADS, st);
    be_pack(st, L0, B, x_api, xw);
    be_zero(st, d_cv, (size_t)B * L0.Ntot * 12 * sizeof(double));
    be_zero(st, d_fv, (size_t)B * L0.Ftot * 4 * sizeof(double));
    be_zero(st, d_bv, (size_t)B * L0.Ftot * 4 * sizeof(double));
    be_grads(st, L0, B, P_coef(pl, persist, 0), rhs, cv0, fv0, bv0, xw, dz_wave, d_coeffs, d_rhs, d_iv, d_cv, d_fv,
             d_bv);
}

extern "C" int pdeop_mg_backward(pdeop_plan* pl, const pdeop_solver_cfg* cfg, const double* rhs, const double* cv0,
                                 const double* fv0, const double* bv0, void* persist, void* scratch, const double* x,
                                 const double* grad_x, double* d_coeffs, double* d_rhs, double* d_iv_rhs, double* d_cv,
                                 double* d_fv, double* d_bv, double* info_out, void* stream) {
    if (check_plan(pl)) return 1;
    if (pl->n_grid < 2) return fail("multigrid path needs n_grid >= 2");
    if (check_cfg(cfg)) return 1;
    Scratch sc = carve(pl, scratch, cfg->restart);
    be_state_reset(stream, sc.state);
    be_pack(stream, pl->lev[0].dev, pl->B, grad_x, sc.atb);
    fgmres(pl, cfg, persist, sc, sc.atb, sc.x, stream);   // dz, same operator and preconditioner (:95)
    if (info_out) be_fg_info(stream, sc.state, info_out);
    run_grads(pl, persist, sc, rhs, cv0, fv0, bv0, x, sc.x, d_coeffs, d_rhs, d_iv_rhs, d_cv, d_fv, d_bv, stream);
    return check_backend();
}

extern "C" int pdeop_mg_forward_converged(pdeop_plan* pl, const pdeop_pcg_cfg* cfg, const double* coeffs,
                                          const double* rhs, const double* iv_rhs, const double* const* cv,
                                          const double* const* fv, const double* const* bv, void* persist,
                                          void* scratch, double* x_out, double* info_out, void* stream) {
    if (check_plan(pl)) return 1;
    if (pl->n_grid < 2) return fail("multigrid path needs n_grid >= 2");
    if (check_pcg(cfg)) return 1;
    Scratch sc = carve(pl, scratch, std::max(5, pl->n_grid + 1));
    setup_operator(pl, coeffs, cv, fv, bv, persist, sc, stream);
    be_atb(stream, pl->lev[0].dev, pl->B, P_coef(pl, persist, 0), rhs, iv_rhs, sc.atb);
    pcg(pl, cfg, persist, sc, sc.atb, sc.x, stream);
    be_unpack(stream, pl->lev[0].dev, pl->B, sc.x, x_out);
    if (info_out) be_fg_info(stream, sc.state, info_out);
    return check_backend();
}

extern "C" int pdeop_mg_backward_converged(pdeop_plan* pl, const pdeop_pcg_cfg* cfg, const double* rhs, const double* cv0,
                                           const double* fv0, const double* bv0, void* persist, void* scratch,
                                           const double* x, const double* grad_x, double* d_coeffs, double* d_rhs,
                                           double* d_iv_rhs, double* d_cv, double* d_fv, double* d_bv, double* info_out,
                                           void* stream) {
    if (check_plan(pl)) return 1;
    if (pl->n_grid < 2) return fail("multigrid path needs n_grid >= 2");
    if (check_pcg(cfg)) return 1;
    Scratch sc = carve(pl, scratch, std::max(5, pl->n_grid + 1));
    be_state_reset(stream, sc.state);
    be_pack(stream, pl->lev[0].dev, pl->B, grad_x, sc.atb);
    pcg(pl, cfg, persist, sc, sc.atb, sc.x, stream);
    if (info_out) be_fg_info(stream, sc.state, info_out);
    run_grads(pl, persist, sc, rhs, cv0, fv0, bv0, x, sc.x, d_coeffs, d_rhs, d_iv_rhs, d_cv, d_fv, d_bv, stream);
    return check_backend();
}

extern "C" int pdeop_dense_forward(pdeop_plan* pl, const double* coeffs, const double* rhs, const double* iv_rhs,
                                   const double* cv0, const double* fv0, const double* bv0, void* persist,
                                   void* scratch, double* x_out, double* info_out, void* stream) {
    if (check_plan(pl)) return 1;
    if (pl->n_grid != 1) return fail("dense path needs a single-level plan");
    Scratch sc = carve(pl, scratch, 1);
    const double* cvp[1] = {cv0};
    const double* fvp[1] = {fv0};
    const double* bvp[1] = {bv0};
    setup_operator(pl, coeffs, cvp, fvp, bvp, persist, sc, stream);
    be_atb(stream, pl->lev[0].dev, pl->B, P_coef(pl, persist, 0), rhs, iv_rhs, sc.atb);
    be_chol_solve(stream, pl->lev[0].dev, pl->B, P_Kd(pl, persist, sc), P_Linv(pl, persist), sc.atb, sc.x, sc.cwork, nullptr, pl->use_chain);
    be_unpack(stream, pl->lev[0].dev, pl->B, sc.x, x_out);
    if (info_out) be_fg_info(stream, sc.state, info_out);
    return check_backend();
}

extern "C" int pdeop_dense_backward(pdeop_plan* pl, const double* rhs, const double* cv0, const double* fv0,
                                    const double* bv0, void* persist, void* scratch, const double* x,
                                    const double* grad_x, double* d_coeffs, double* d_rhs, double* d_iv_rhs,
                                    double* d_cv, double* d_fv, double* d_bv, double* info_out, void* stream) {
    if (check_plan(pl)) return 1;
    if (pl->n_grid != 1) return fail("dense path needs a single-level plan");
    Scratch sc = carve(pl, scratch, 1);
    be_state_reset(stream, sc.state);
    be_pack(stream, pl->lev[0].dev, pl->B, grad_x, sc.atb);
    be_chol_solve(stream, pl->lev[0].dev, pl->B, P_Kd(pl, persist, sc), P_Linv(pl, persist), sc.atb, sc.x, sc.cwork, nullptr, pl->use_chain);  // dz (:65)
    if (info_out) be_fg_info(stream, sc.state, info_out);
    run_grads(pl, persist, sc, rhs, cv0, fv0, bv0, x, sc.x, d_coeffs, d_rhs, d_iv_rhs, d_cv, d_fv, d_bv, stream);
    return check_backend();
}

extern "C" int pdeop_fgmres(pdeop_plan* pl, const pdeop_solver_cfg* cfg, int back, const double* b, double* x_out,
                            double* info_out, double* hess_out, void* persist, void* scratch, void* stream) {
    (void)back;
    if (check_plan(pl)) return 1;
    if (check_cfg(cfg)) return 1;
    Scratch sc = carve(pl, scratch, cfg->restart);
    be_state_reset(stream, sc.state);
    be_pack(stream, pl->lev[0].dev, pl->B, b, sc.atb);
    fgmres(pl, cfg, persist, sc, sc.atb, sc.x, stream);
    be_unpack(stream, pl->lev[0].dev, pl->B, sc.x, x_out);
    if (info_out) be_fg_info(stream, sc.state, info_out);
    if (hess_out) be_fg_hess(stream, sc.state, cfg->restart, hess_out);
    return check_backend();
}

extern "C" int pdeop_stage(pdeop_plan* pl, const pdeop_solver_cfg* cfg, int stage, int level, int count,
                           const double* in1, const double* in2, double* out, void* persist, void* scratch,
                           void* stream) {
    if (check_plan(pl)) return 1;
    if (check_cfg(cfg)) return 1;
    if (level < 0 || level >= pl->n_grid) return fail("level out of range");
    Scratch sc = carve(pl, scratch, cfg->restart);
    const int B = pl->B;
    const LevelDev& L = pl->lev[level].dev;
    // level-0-sized temporaries are large enough for any level
    double* t1 = sc.V;
    double* t2 = sc.V + sc.n0;
    double* t3 = sc.Z;
    be_state_reset(stream, sc.state);
    switch (stage) {
        case PDEOP_STAGE_APPLY_K:
            be_pack(stream, L, B, in1, t1);
            be_apply_k(stream, L, B, P_T(pl, persist, level), P_coef(pl, persist, level), t1, nullptr, t2, 0, nullptr);
            be_unpack(stream, L, B, t2, out);
            break;
        case PDEOP_STAGE_GS:
            be_pack(stream, L, B, in1, t1);
            be_pack(stream, L, B, in2, t2);
            be_gs(stream, L, B, P_T(pl, persist, level), P_coef(pl, persist, level), P_dinv(pl, persist, level), t1, t2,
                  sc.gs_stash, sc.gs_stash_stride, count, nullptr, cfg->gs_variant, pl->gs_pipe);
            be_unpack(stream, L, B, t2, out);
            break;
        case PDEOP_STAGE_RESTRICT: {
            if (level + 1 >= pl->n_grid) return fail("no coarser level");
            const LevelDev& Lc = pl->lev[level + 1].dev;
            be_pack(stream, L, B, in1, t1);
            be_interp(stream, L, Lc, B, L.M, t1, t2, 0, nullptr);
            be_unpack(stream, Lc, B, t2, out);
            break;
        }
        case PDEOP_STAGE_PROLONG: {
            if (level < 1) return fail("no finer level");
            const LevelDev& Lf = pl->lev[level - 1].dev;
            be_pack(stream, L, B, in1, t1);
            be_interp(stream, L, Lf, B, L.M, t1, t2, 0, nullptr);
            be_unpack(stream, Lf, B, t2, out);
            break;
        }
        case PDEOP_STAGE_VCYCLE:
            if (level != 0) return fail("V-cycle stage starts at level 0");
            if (pl->n_grid < 2) return fail("V-cycle needs n_grid >= 2");
            be_pack(stream, L, B, in1, t1);
            vcycle_start(pl, cfg, persist, sc, t1, t3, sc.w, stream);
            be_unpack(stream, L, B, t3, out);
            break;
        case PDEOP_STAGE_COARSE_SOLVE:
            if (level != pl->n_grid - 1) return fail("coarse solve runs on the last level");
            be_pack(stream, L, B, in1, t1);
            be_chol_solve(stream, L, B, P_Kd(pl, persist, sc), P_Linv(pl, persist), t1, t2, sc.cwork, nullptr, pl->use_chain);
            be_unpack(stream, L, B, t2, out);
            break;
        case PDEOP_STAGE_ATB:
            if (level != 0) return fail("A^T b is defined on level 0");
            be_atb(stream, L, B, P_coef(pl, persist, 0), in1, in2, t1);
            be_unpack(stream, L, B, t1, out);
            break;
        default:
            return fail("unknown stage");
    }
    return check_backend();
}
