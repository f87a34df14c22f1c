/* Plain-C restatement of the reference's process model -- ORACLE / CPU BASELINE ONLY.
 *
 * Test infrastructure: linked only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs.  Parity status: unpinned against Pyomo/CONOPT
 * (see oracle/__init__.py); cross-checked against oracle/kinetics.py and kinetics_mp.py.
 *
 * Equations: reference run_twostage.py:49-55 (mass balances, dcdt/t_f == rhs),
 * :22-23 (du/dtau = F_in/t_f), :62-65 (g1, g2), :69-74 (big-M constants, initial state),
 * :76-79 (Radau collocation nfe=8, ncp=3, F_in piecewise constant per element).
 * Reduction: P_i = sum_j w_j 1[g1<=0 and g2<=0]  (DEUS efp).
 *
 * d columns: (t_f, T[, cA0[, Fbar[, pB, pA]]]); p columns: (E1, E2, k1, k2).
 * Each (d,theta) pair: 8 finite elements, full Newton on the 9x9 collocation system,
 * dense LU with partial pivoting.  OpenMP over design points.
 *
 * Two formulations of the SAME collocation equations (oracle_set_formulation):
 *   0 "dense9"   the 9 unknowns of an element (cA, cB, cC at three points) in one pivoted Newton system -- what
 *                SURVEY.md A.3 / A.5 count (about 22 kflop per evaluation);
 *   1 "blocktri" what the GPU kernels solve: the system is block lower triangular (A -> B -> C), so cA is a 3-unknown
 *                Newton solve, cB one linear 3x3 solve given cA, cC a quadrature (about 3 kflop per evaluation).
 *                Same Newton iterates for cA, hence the same solution; lets bench.py separate what the hardware
 *                buys from what the restructuring buys.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NFE 8
#define NCP 3
#define NS 3
#define NN 9

/* adot[k][j] = l_k'(tau_j), Radau nodes {0,(4-sqrt6)/10,(4+sqrt6)/10,1}; written as closed
 * forms so the compiler rounds them once to fp64 (values agree with oracle/radau.py). */
static double A[4][4];
static double BW[3];
static int tables_ready = 0;

static void init_tables(void)
{
    if (tables_ready) return;
    const double s6 = sqrt(6.0);
    const double tau[4] = {0.0, (4.0 - s6) / 10.0, (4.0 + s6) / 10.0, 1.0};
    for (int k = 0; k < 4; ++k)
        for (int j = 0; j < 4; ++j) {
            long double s = 0.0L;
            for (int m = 0; m < 4; ++m) {
                if (m == k) continue;
                long double p = 1.0L / ((long double)tau[k] - tau[m]);
                for (int l = 0; l < 4; ++l) {
                    if (l == k || l == m) continue;
                    p *= ((long double)tau[j] - tau[l]) / ((long double)tau[k] - tau[l]);
                }
                s += p;
            }
            A[k][j] = (double)s;
        }
    /* exact values override the numerically built ones where they are rationals */
    A[0][0] = -9.0; A[0][3] = -3.0; A[3][0] = 1.0 / 3.0; A[3][3] = 5.0;
    BW[0] = (16.0 - s6) / 36.0; BW[1] = (16.0 + s6) / 36.0; BW[2] = 1.0 / 9.0;
    tables_ready = 1;
}

/* Optional: load the exact fp64 tables oracle/radau.py computed with mpmath. */
void oracle_set_tables(const double* adot16, const double* bw3)
{
    for (int k = 0; k < 4; ++k) for (int j = 0; j < 4; ++j) A[k][j] = adot16[4 * k + j];
    for (int j = 0; j < 3; ++j) BW[j] = bw3[j];
    tables_ready = 1;
}

/* Number of OpenMP threads for the grid evaluation (torchrun exports OMP_NUM_THREADS=1). */
int oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

static int lu_solve9(double M[NN][NN], double r[NN])
{
    for (int k = 0; k < NN; ++k) {
        int piv = k; double best = fabs(M[k][k]);
        for (int i = k + 1; i < NN; ++i) if (fabs(M[i][k]) > best) { best = fabs(M[i][k]); piv = i; }
        if (best == 0.0) return -1;
        if (piv != k) {
            for (int c = 0; c < NN; ++c) { double t = M[k][c]; M[k][c] = M[piv][c]; M[piv][c] = t; }
            double t = r[k]; r[k] = r[piv]; r[piv] = t;
        }
        const double inv = 1.0 / M[k][k];
        for (int i = k + 1; i < NN; ++i) {
            const double l = M[i][k] * inv;
            if (l == 0.0) continue;
            for (int c = k + 1; c < NN; ++c) M[i][c] -= l * M[k][c];
            r[i] -= l * r[k];
        }
    }
    for (int k = NN - 1; k >= 0; --k) {
        double s = r[k];
        for (int c = k + 1; c < NN; ++c) s -= M[k][c] * r[c];
        r[k] = s / M[k][k];
    }
    return 0;
}

/* One (d,theta) evaluation.  Returns total Newton iterations, or -1 on failure. */
static int eval_one(double tf, double T, double cA0, const double* F, double pB, double pA,
                    double E1, double E2, double k1, double k2, double tol, int max_it,
                    double* g1, double* g2)
{
    const double K1 = k1 * exp(-E1 / T), K2 = k2 * exp(-E2 / T);
    const double h = 1.0 / NFE;
    double x0[NS] = {cA0, 0.0, 0.0};
    double u = 0.0;
    int total = 0;
    for (int e = 0; e < NFE; ++e) {
        double X[NCP][NS];
        for (int j = 0; j < NCP; ++j) for (int n = 0; n < NS; ++n) X[j][n] = x0[n];
        int it = 0, conv = 0;
        while (it < max_it && !conv) {
            double M[NN][NN]; double r[NN];
            memset(M, 0, sizeof M);
            for (int j = 0; j < NCP; ++j) {
                const double cA = X[j][0], cB = X[j][1];
                const double r1 = K1 * cA * cA, r2 = K2 * cB;
                const double f[NS] = {tf * (-2.0 * r1 + F[e]), tf * (r1 - r2), tf * r2};
                const double d1 = 2.0 * K1 * cA;
                for (int n = 0; n < NS; ++n) {
                    double lin = A[0][j + 1] * x0[n];
                    for (int k = 0; k < NCP; ++k) {
                        lin += A[k + 1][j + 1] * X[k][n];
                        M[NS * j + n][NS * k + n] += A[k + 1][j + 1];
                    }
                    r[NS * j + n] = -(lin - h * f[n]);
                }
                M[NS * j + 0][NS * j + 0] -= h * (-2.0 * tf * d1);
                M[NS * j + 1][NS * j + 0] -= h * (tf * d1);
                M[NS * j + 1][NS * j + 1] -= h * (-tf * K2);
                M[NS * j + 2][NS * j + 1] -= h * (tf * K2);
            }
            if (lu_solve9(M, r) != 0) return -1;
            double dmax = 0.0, xmax = 1.0;
            for (int j = 0; j < NCP; ++j) for (int n = 0; n < NS; ++n) {
                X[j][n] += r[NS * j + n];
                if (fabs(r[NS * j + n]) > dmax) dmax = fabs(r[NS * j + n]);
                if (fabs(X[j][n]) > xmax) xmax = fabs(X[j][n]);
            }
            ++it;
            if (!(dmax == dmax)) return -1;
            conv = dmax <= tol * xmax;
        }
        if (!conv) return -1;
        total += it;
        u += h * F[e] * (BW[0] + BW[1] + BW[2]) / tf;
        for (int n = 0; n < NS; ++n) x0[n] = X[NCP - 1][n];
    }
    const double cA = x0[0], cB = x0[1], cC = x0[2];
    *g1 = 0.8 - (cB + 1e-2) / (cA + cB + cC + 1e-2);
    *g2 = -(pB * cB - pA * (cA0 + u * tf) - 128.0 * (tf + 30.0));
    return total;
}

static int lu_solve3(double M[3][3], double r[3])
{
    for (int k = 0; k < 3; ++k) {
        int piv = k; double best = fabs(M[k][k]);
        for (int i = k + 1; i < 3; ++i) if (fabs(M[i][k]) > best) { best = fabs(M[i][k]); piv = i; }
        if (best == 0.0) return -1;
        if (piv != k) {
            for (int c = 0; c < 3; ++c) { double t = M[k][c]; M[k][c] = M[piv][c]; M[piv][c] = t; }
            double t = r[k]; r[k] = r[piv]; r[piv] = t;
        }
        const double inv = 1.0 / M[k][k];
        for (int i = k + 1; i < 3; ++i) {
            const double l = M[i][k] * inv;
            for (int c = k + 1; c < 3; ++c) M[i][c] -= l * M[k][c];
            r[i] -= l * r[k];
        }
    }
    for (int k = 2; k >= 0; --k) {
        double s = r[k];
        for (int c = k + 1; c < 3; ++c) s -= M[k][c] * r[c];
        r[k] = s / M[k][k];
    }
    return 0;
}

/* The block-triangular formulation of the same equations: cA by Newton (3 unknowns), cB by one linear solve,
 * cC by its quadrature  cC(end) = cC(start) + h sum_j bw_j t_f K2 cB_j  (Radau IIA weights: exact for the collocation
 * polynomial).  Returns the Newton iterations of the cA blocks, or -1. */
static int eval_one_blocktri(double tf, double T, double cA0, const double* F, double pB, double pA,
                             double E1, double E2, double k1, double k2, double tol, int max_it,
                             double* g1, double* g2)
{
    const double K1 = k1 * exp(-E1 / T), K2 = k2 * exp(-E2 / T);
    const double h = 1.0 / NFE;
    double cA = cA0, cB = 0.0, cC = 0.0, u = 0.0;
    int total = 0;
    for (int e = 0; e < NFE; ++e) {
        double XA[NCP] = {cA, cA, cA};
        int it = 0, conv = 0;
        while (it < max_it && !conv) {
            double M[3][3], r[3];
            for (int j = 0; j < NCP; ++j) {
                double lin = A[0][j + 1] * cA;
                for (int k = 0; k < NCP; ++k) { lin += A[k + 1][j + 1] * XA[k]; M[j][k] = A[k + 1][j + 1]; }
                const double f = tf * (-2.0 * K1 * XA[j] * XA[j] + F[e]);
                r[j] = -(lin - h * f);
                M[j][j] -= h * (-4.0 * tf * K1 * XA[j]);
            }
            if (lu_solve3(M, r) != 0) return -1;
            double dmax = 0.0, xmax = 1.0;
            for (int j = 0; j < NCP; ++j) {
                XA[j] += r[j];
                if (fabs(r[j]) > dmax) dmax = fabs(r[j]);
                if (fabs(XA[j]) > xmax) xmax = fabs(XA[j]);
            }
            ++it;
            if (!(dmax == dmax)) return -1;
            conv = dmax <= tol * xmax;
        }
        if (!conv) return -1;
        total += it;
        {
            double M[3][3], r[3];
            for (int j = 0; j < NCP; ++j) {
                for (int k = 0; k < NCP; ++k) M[j][k] = A[k + 1][j + 1];
                M[j][j] += h * tf * K2;
                r[j] = h * tf * K1 * XA[j] * XA[j] - A[0][j + 1] * cB;
            }
            if (lu_solve3(M, r) != 0) return -1;
            double q = 0.0;
            for (int j = 0; j < NCP; ++j) q += BW[j] * r[j];
            cC += h * tf * K2 * q;
            cB = r[NCP - 1];
            if (!(cB == cB)) return -1;
        }
        u += h * F[e] * (BW[0] + BW[1] + BW[2]) / tf;
        cA = XA[NCP - 1];
    }
    *g1 = 0.8 - (cB + 1e-2) / (cA + cB + cC + 1e-2);
    *g2 = -(pB * cB - pA * (cA0 + u * tf) - 128.0 * (tf + 30.0));
    return total;
}

static int formulation = 0;
/* 0: dense 9x9 Newton per element (default), 1: block-triangular 3 + 3.  Returns the previous value. */
int oracle_set_formulation(int f)
{
    const int old = formulation;
    if (f == 0 || f == 1) formulation = f;
    return old;
}

/* Full grid.  F: NULL (zeros), or (f_rows, 8) with f_rows in {1, n_d}; ignored when d_dim >= 4.
 * g_out: NULL or (n_d, n_p, 2).  P_out: NULL or (n_d).  Returns total Newton iterations
 * (negative count of failed scenarios if any failed; failed scenarios count as infeasible). */
long long oracle_kinetics_grid(long long n_d, int d_dim, const double* d, long long n_p, const double* p,
                               const double* w, const double* F, long long f_rows, double tol, int max_it,
                               double* g_out, double* P_out)
{
    init_tables();
    long long iters = 0, fails = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : iters, fails)
    for (long long i = 0; i < n_d; ++i) {
        const double* di = d + i * d_dim;
        const double tf = di[0], T = di[1];
        const double cA0 = d_dim >= 3 ? di[2] : 2000.0;
        double Fl[NFE];
        for (int e = 0; e < NFE; ++e)
            Fl[e] = d_dim >= 4 ? di[3] : (F ? F[(f_rows > 1 ? i : 0) * NFE + e] : 0.0);
        const double pB = d_dim >= 6 ? di[4] : 100.0, pA = d_dim >= 6 ? di[5] : 20.0;
        double acc = 0.0;
        for (long long j = 0; j < n_p; ++j) {
            const double* pj = p + j * 4;
            double g1, g2;
            int it = formulation ? eval_one_blocktri(tf, T, cA0, Fl, pB, pA, pj[0], pj[1], pj[2], pj[3], tol, max_it, &g1, &g2)
                                 : eval_one(tf, T, cA0, Fl, pB, pA, pj[0], pj[1], pj[2], pj[3], tol, max_it, &g1, &g2);
            if (it < 0) { ++fails; g1 = g2 = NAN; } else iters += it;
            if (g_out) { g_out[(i * n_p + j) * 2] = g1; g_out[(i * n_p + j) * 2 + 1] = g2; }
            if (it >= 0 && g1 <= 0.0 && g2 <= 0.0) acc += w ? w[j] : 1.0;
        }
        if (P_out) P_out[i] = acc;
    }
    return fails ? -fails : iters;
}
