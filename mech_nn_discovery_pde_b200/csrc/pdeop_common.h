// pdeop -- B200-native differentiable PDE-layer solve.  Shared definitions.
//
// Internal data model:
//  * Every grid is embedded in 3 internal axes (N0,N1,N2); a d-dimensional problem uses the last d
//    axes ("active" axes), leading axes have extent 1.
//  * Grid points are stored in WAVE ORDER: sorted by hyperplane s=i0+i1+i2, then (i0,i1).  Points of
//    one hyperplane are independent under lexicographic Gauss-Seidel (the normal operator couples
//    points only along grid axes), so a hyperplane is a contiguous, coalesced index range.
//  * Vectors are channel-planar: v[b][m][w], m in [0,M), M=1+2d channels (u, u_c, u_cc), w wave index.
//  * The constraint matrix A is never formed.  Per instance and axis a small table T holds the
//    axis part of K=A^T A per line position (30 numbers), built from the per-line row values.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PDEOP_HD __host__ __device__ __forceinline__
#else
#define PDEOP_HD inline
#endif

namespace pdeop {

constexpr int kTabEntries = 30;   // Cuu[9] | Cup[9] | Cuq[9] | pp | qq | pq
constexpr int kTabPad = 4;        // stencil radius of K along an axis
// Fixed table pitch (max extent 1023 + 2*4 padding, rounded): a compile-time constant so that every table
// load of a stencil kernel is [position pointer + immediate offset] with no address arithmetic.
constexpr int kTabPitch = 1032;
// Hyperplane lag between pipelined Gauss-Seidel sweeps: radius + 2.  Radius + 1 keeps the sweeps independent; one
// more lets the gather of step t+1 (everything but the backward distance-1 neighbours) run during step t, because
// the farthest forward neighbour (s+1+4) was then finished by the previous sweep in step t-1, not in step t.
constexpr int kGsLag = 6;
// The unsplit kernel gathers and finishes a point in the same step: radius + 1 suffices (4 steps fewer per call).
constexpr int kGsLagUnsplit = 5;
constexpr int kMaxRestart = 32;
// Block size of the dense triangular solves (inverse diagonal blocks).  128: the chain solver prefetches, per CTA,
// the previous block's columns of its rows (kSolveBlk/4 rows x kSolveBlk columns) into shared memory; at 128 that
// buffer is 32 KB and leaves >160 KB for the TMA stages that keep HBM busy, at 256 it would take 128 KB.
constexpr int kSolveBlk = 128;

enum TabIdx { T_UU = 0, T_UP = 9, T_UQ = 18, T_PP = 27, T_QQ = 28, T_PQ = 29 };

// Immutable per-level index tables (device pointers in the CUDA build, host pointers in the emulator).
struct LevelDev {
    int N[3];        // internal extents
    int D;           // active axes (problem dimension); active internal axes are 3-D .. 2
    int M;           // channels = 1 + 2*D
    int G;           // grid points
    int S;           // hyperplanes = N0+N1+N2-2
    int P;           // table pitch (= kTabPitch)
    int n_init;      // initial/boundary rows
    int n_eq;        // equation rows
    int Ntot;        // sum of active extents          (central row-value table length)
    int Ftot;        // sum of (active extents - 1)    (forward/backward row-value table length)
    int cvoff[3];    // per active axis a: offset into the central line table
    int fvoff[3];    // per active axis a: offset into the forward/backward line tables
    const int* coord;    // [G]  i0 | i1<<10 | i2<<20 | eq<<30 (eq: carries an equation row), wave order
    const int* flags;    // [G]  bit0: carries an equation row; bits 4+2m..5+2m: # initial rows on channel m
    const int* hstart;   // [S+1] first wave index of each hyperplane
    const int* rowbase;  // [(S+8)*N0] pos(i0,i1,i2) = rowbase[(s+4)*N0+i0] + i1  (4 spare ints either side)
    const int* init_w;   // [n_init] wave index of each initial row's variable
    const int* init_m;   // [n_init] channel of each initial row's variable
    // Dense (coarsest / dense-layer) ordering: unknown (w,m) -> band[w]*M + m, where band[] numbers the grid
    // points with the LONGEST axis outermost, so that K is banded with half-bandwidth bw (stencil radius 4
    // along the outer axis => bw = 4*inner*M + M - 1).  Cholesky creates no fill outside the band.
    const int* band;     // [G]
    int bw;
    // Total derivative order of the constraint system (lp_pde_central_diff.py:304-315).  Order 1 has no second-
    // derivative unknowns: it runs on the same (u, u_c, u_cc) layout with every u_cc coupling exactly zero (the
    // caller passes zero second-order rows) and a unit diagonal on the u_cc channels, so those unknowns stay zero and
    // the (u, u_c) block is the order-1 normal operator bit for bit.
    int order;
};

PDEOP_HD void unpack_coord(int c, int& i0, int& i1, int& i2) {
    i0 = c & 1023;
    i1 = (c >> 10) & 1023;
    i2 = (c >> 20) & 1023;
}

PDEOP_HD bool coord_eq(int c) { return (c >> 30) & 1; }

PDEOP_HD int wave_pos(const LevelDev& L, int i0, int i1, int i2) {
    return L.rowbase[(i0 + i1 + i2 + 4) * L.N[0] + i0] + i1;
}

PDEOP_HD int nat_index(const LevelDev& L, int i0, int i1, int i2) {
    return (i0 * L.N[1] + i1) * L.N[2] + i2;
}

// Device-resident FGMRES bookkeeping: nothing here is read by the host inside a solve.
struct FgmresState {
    double H[(kMaxRestart + 1) * kMaxRestart];  // (restart+1) x restart, row-major with ld = restart
    double e[kMaxRestart + 1];
    double y[kMaxRestart];
    double bnorm;
    double rnorm;
    double tmp[kMaxRestart + 2];
    int iters;
    int done;
    int chol_info;   // 0 ok, k>0: pivot k of some instance was not positive
    int pad;
};

// Scratch area for two-phase reductions, placed right after the FgmresState in the scratch buffer:
// two ping-pong areas of kReduceBlocks x (kMaxRestart+1) block partials.
constexpr int kReduceBlocks = 592;   // 4 x 148 SMs
constexpr size_t kStateBytes = (sizeof(FgmresState) + 255) / 256 * 256;
constexpr size_t kPartialDoubles = (size_t)kReduceBlocks * (kMaxRestart + 1);
constexpr size_t kStateAreaDoubles = kStateBytes / 8 + 2 * kPartialDoubles;

PDEOP_HD double* state_partials(const FgmresState* s, int which) {
    return (double*)((char*)s + kStateBytes) + (size_t)which * kPartialDoubles;
}

}  // namespace pdeop
