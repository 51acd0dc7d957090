// pdeop -- least-squares solve of the (m+1) x m FGMRES Hessenberg system  min || H y - e ||  with
// e = (rnorm, 0, ..., 0)  (fgmres.py:166, torch.linalg.lstsq in the reference).  H is upper Hessenberg,
// so m Givens rotations reduce it to triangular form; one thread does it on the device.
#pragma once
#include "pdeop_common.h"

namespace pdeop {

// H: (m+1) x m row-major with leading dimension m (left untouched); e: m+1 entries; y: m outputs.
PDEOP_HD void hessenberg_lstsq(const double* H, const double* e, int m, double* y) {
    double R[(kMaxRestart + 1) * kMaxRestart];
    double g[kMaxRestart + 1];
    for (int i = 0; i <= m; ++i) {
        g[i] = e[i];
        for (int j = 0; j < m; ++j) R[i * m + j] = H[i * m + j];
    }
    for (int j = 0; j < m; ++j) {
        const double a = R[j * m + j], b = R[(j + 1) * m + j];
        const double r = hypot(a, b);
        double c = 1.0, s = 0.0;
        if (r != 0.0) {
            c = a / r;
            s = b / r;
        }
        for (int k = j; k < m; ++k) {
            const double t1 = R[j * m + k], t2 = R[(j + 1) * m + k];
            R[j * m + k] = c * t1 + s * t2;
            R[(j + 1) * m + k] = -s * t1 + c * t2;
        }
        const double g1 = g[j], g2 = g[j + 1];
        g[j] = c * g1 + s * g2;
        g[j + 1] = -s * g1 + c * g2;
    }
    for (int i = m - 1; i >= 0; --i) {
        double v = g[i];
        for (int k = i + 1; k < m; ++k) v -= R[i * m + k] * y[k];
        y[i] = v / R[i * m + i];
    }
}

}  // namespace pdeop
