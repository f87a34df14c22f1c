// X_PCHAIN: the whole finite-element loop of a chain of polynomial one-state blocks in ONE instruction.
//
// When every block of the block-triangular state solve is a one-state block with  f_b = c0 + c1 x_b + c2 x_b^2  and every
// coefficient / quadrature integrand is a single term
//
//        (narrow value) [x (control of this finite element)] [x (state of an EARLIER block)^p,  p = 1, 2]
//
// (mass-action kinetics: A -> B -> C of run_twostage.py:49-55 is  f_A = t_f F_in - 2 t_f K1 cA^2,
//  f_B = t_f K1 cA^2 - t_f K2 cB,  dcC = t_f K2 cB,  du = F_in / t_f), the assembler (assembler.h, try_chain) replaces
// the element loop -- [LDCTRL, elem ops,] POLY1 ..., ELEM_END, 3 dispatches per element and a round trip of every state
// through the shared-memory register file per block -- by this one instruction: the states and quadratures stay in
// registers over all nfe elements, nothing is dispatched, and nothing but the narrow coefficients is loaded.
//
// Same equations and same start (x_j = x(element start)) as POLY1 (feas_kernel.cuh); the stop rule is POLY1's with the
// relative scale |x|^2 taken at the element start (3 x_s^2, one limit per element) instead of at every iterate.  Algebraic
// savings:
//   * coefficients are pre-scaled by h once per element and the x(element start) column of the differentiation matrix is
//     folded into the constant term; the diagonal of the differentiation matrix joins the Horner form, so a residual is
//     4 FMAs per point and the Jacobian diagonal 1 (the Newton matrix is assembled negated:  N = h J - A,  x -= N^-1 r);
//   * a block that is LINEAR with an element- and point-invariant c1 (first-order decay, -t_f K2 cB) has the same
//     collocation matrix  A - h c1 I  on every element: it is inverted once per scenario (9 values in register-file
//     units that are dead during the loop) and every element costs one 3x3 matrix-vector product instead of an LU.
//
// Instruction layout (uint4 units):
//   w0 = { X_PCHAIN, nb | nq << 8, max_it, 0 }       w1 = { 0, 0, 0, 0 }
//   per block b (3 words):  { x_off, flops per iteration, kind, minv_off }
//                           { c0.n_off, c0.meta, c1.n_off, c1.meta }      { c2.n_off, c2.meta, 0, 0 }
//   per quadrature (1):     { q_off, n_off, meta, 0 }
//   meta = present | has_n << 1 | has_ctrl << 2 | negated << 3 | ctrl << 4 | p << 8 | source block << 12
// The kernel does not decode this from the instruction stream: the host copies it into the fixed constant-bank struct
// c_chain (ChainDesc), so every field is a direct constant operand and every branch on it a uniform one -- read through
// a program counter the descriptors arrived in vector registers (LDC R, c[3][R]) and the per-element decode cost as much
// as the Newton iterations (profiles/r02_v1_*).  A chain program has a fixed shape,
//     scalar ops (pro tape, INIT, element-invariant control loads + elem tape) | PCHAIN | XEND | scalar ops (g tape) | END,
// so the chain instantiation of feas_kernel runs two counted loops around poly_chain instead of a dispatch loop; INIT and
// XEND are not executed either: the chain reads the initial values where they are (constant pool or parameter slot) and
// writes the end values to the slots the g tape reads.
#pragma once
#include "xvm_ops.cuh"

namespace pydsb {

constexpr int CHAIN_MAXB = 3;              // blocks in one chain
constexpr int CHAIN_MAXQ = 4;              // quadratures advanced by the chain
constexpr unsigned CT_PRESENT = 1u, CT_HAS_N = 2u, CT_HAS_CTRL = 4u, CT_NEG = 8u;
constexpr unsigned CHAIN_NEWTON = 0, CHAIN_LINEAR = 1, CHAIN_LINCONST = 2;
// algorithmic flops (SURVEY 8d convention): one-off inverse = LU (2/3 n^3) + three solves (2 n^2 each);
// one element of a LINCONST block = right-hand side (2 n) + matrix-vector product (2 n^2)
constexpr unsigned CHAIN_INV_FLOPS = 18 + 3 * 18, CHAIN_LINCONST_FLOPS = 6 + 18;

__host__ __device__ constexpr unsigned chain_meta(bool has_n, bool has_ctrl, unsigned ctrl, unsigned p, unsigned src)
{
    return CT_PRESENT | (has_n ? CT_HAS_N : 0u) | (has_ctrl ? CT_HAS_CTRL : 0u) | (ctrl << 4) | (p << 8) | (src << 12);
}
__host__ __device__ constexpr int chain_length(int nb, int nq) { return 2 + 3 * nb + nq; }

struct ChainTermD { unsigned n_off, meta; };
// x0: where the initial value comes from (byte offset; x0_const: in the constant pool, else a slot); xe_off: the slot the
// g tape reads the end value from
struct ChainBlockD { unsigned x_off, flops, kind, minv_off; ChainTermD c[3]; unsigned x0_off, x0_const, xe_off, pad; };
struct ChainQuadD { unsigned q_off, n_off, meta, q0_off, q0_const, pad[3]; };
struct ChainDesc {
    unsigned nb, nq, max_it, pre_end;          // scalar ops [0, pre_end) run before the chain ...
    unsigned g_begin, g_end, n_par, ng;        // ... and [g_begin, g_end) after it (uint4 indices into c_prog)
    ChainBlockD blk[CHAIN_MAXB];
    ChainQuadD quad[CHAIN_MAXQ];
    unsigned pre_run0, g_run0, pad2, pad3;     // FMAG ops at the head of either scalar section (later runs: w1.w of the op before)
    unsigned par_off[16];                      // slot of every parameter (d first, then theta)
    unsigned g_off[8];                         // slot of every constraint value
    unsigned z_off[16];                        // where the initial / fixed value of control value zi lives ...
    unsigned z_const, nz, pad4, pad5;          // ... bit zi of z_const: in the constant pool, else a slot
};
__constant__ ChainDesc c_chain;

// Compile-time shape of a chain.  The structure of a chain -- how many blocks, which kind, which coefficient is present and
// whether it is narrow, narrow * x or narrow * x^2 of which earlier block -- decides which code runs per element; decoded at
// run time (the generic instantiation, SB = 0) that is ~330 instructions per element of selects and uniform branches around
// ~80 of arithmetic.  A shape registered in chain_shapes.h gets an instantiation in which all of it is folded at compile
// time; offsets, iteration cap and the control indices stay in the descriptor.  Bit layout (host: chain_shape_bits):
//   SB  [0:1] nb   [2] some term depends on a control   block b at 3 + 19 b: kind(2) c0(7) c1(7) c2(3)
//   SQ  [0:2] nq   quadrature q at 3 + 7 q: term(7)     term = present | has_n << 1 | p << 2 | source block << 4 | negated << 6
inline unsigned chain_term_bits(unsigned meta)
{
    return (meta & 3u) | (((meta >> 8) & 3u) << 2) | (((meta >> 12) & 3u) << 4) | ((meta & CT_NEG) ? 64u : 0u);
}
inline void chain_shape_bits(const ChainDesc& cd, unsigned long long& sb, unsigned& sq)
{
    sb = cd.nb & 3u;
    sq = cd.nq & 7u;
    bool ctrl = false;
    for (unsigned b = 0; b < cd.nb; ++b) {
        unsigned long long v = cd.blk[b].kind & 3u;
        for (int k = 0; k < 3; ++k) {
            const unsigned t = chain_term_bits(cd.blk[b].c[k].meta);
            // c2 is narrow: present, has_n, negated
            v |= (unsigned long long)(k == 2 ? ((t & 3u) | ((t >> 6) << 2)) : t) << (2 + 7 * k);
            ctrl = ctrl || (cd.blk[b].c[k].meta & CT_HAS_CTRL);
        }
        sb |= v << (3 + 19 * b);
    }
    for (unsigned q = 0; q < cd.nq; ++q) {
        sq |= chain_term_bits(cd.quad[q].meta) << (3 + 7 * q);
        ctrl = ctrl || (cd.quad[q].meta & CT_HAS_CTRL);
    }
    if (ctrl) sb |= 4ull;
}

template <unsigned long long SB, unsigned SQ>
struct ChainShape {
    static constexpr bool generic = SB == 0ull;
    static constexpr bool ctrl = generic || ((SB >> 2) & 1ull);
    static __device__ __forceinline__ unsigned unpack(unsigned t)
    {
        return (t & 3u) | (((t >> 2) & 3u) << 8) | (((t >> 4) & 3u) << 12) | ((t & 64u) ? CT_NEG : 0u);
    }
    static __device__ __forceinline__ int nb() { return generic ? (int)c_chain.nb : (int)(SB & 3ull); }
    static __device__ __forceinline__ int nq() { return generic ? (int)c_chain.nq : (int)(SQ & 7u); }
    static __device__ __forceinline__ unsigned kind(int b)
    {
        return generic ? c_chain.blk[b].kind : (unsigned)((SB >> (3 + 19 * b)) & 3ull);
    }
    // meta word of coefficient k of block b: structural bits from the shape, control bits from the descriptor
    static __device__ __forceinline__ unsigned meta(int b, int k)
    {
        if (generic) return c_chain.blk[b].c[k].meta;
        unsigned t = (unsigned)((SB >> (3 + 19 * b + 2 + 7 * k)) & (k == 2 ? 7ull : 127ull));
        if (k == 2) t = (t & 3u) | ((t >> 2) << 6);
        const unsigned st = unpack(t);
        return ctrl ? (st | (c_chain.blk[b].c[k].meta & (CT_HAS_CTRL | 0xF0u))) : st;
    }
    static __device__ __forceinline__ unsigned qmeta(int q)
    {
        if (generic) return c_chain.quad[q].meta;
        const unsigned st = unpack((SQ >> (3 + 7 * q)) & 127u);
        return ctrl ? (st | (c_chain.quad[q].meta & (CT_HAS_CTRL | 0xF0u))) : st;
    }
};

#define CH_A(jj, kk) img.adot[4 * ((kk) + 1) + (jj) + 1]      /* d l_{kk+1} / dtau at point jj+1 */
#define CH_A0(jj) img.adot[(jj) + 1]                          /* d l_0 / dtau at point jj+1 */

// narrow factor of a term on finite element e, times `scale`
template <int SPT, bool CTRL, class CtrlLd>
__device__ __forceinline__ void chain_narrow(unsigned n_off, unsigned meta, int e, const char* __restrict__ base,
                                             CtrlLd& ctrl_ld, double scale, double (&nv)[SPT])
{
    typedef Vec<SPT> VecT;
    if (meta & CT_NEG) scale = -scale;
    if (meta & CT_HAS_N) {
        const VecT v = LDV(n_off);
#pragma unroll
        for (int s = 0; s < SPT; ++s) nv[s] = scale * v.v[s];
    } else {
#pragma unroll
        for (int s = 0; s < SPT; ++s) nv[s] = scale;
    }
    if (CTRL && (meta & CT_HAS_CTRL)) {
        const int c = (int)((meta >> 4) & 0xFu);
#pragma unroll
        for (int s = 0; s < SPT; ++s) nv[s] *= ctrl_ld(e, c, s);
    }
}

// (state of block `sb`)^p at point j of scenario s; nsrc = blocks that may be referenced (a constant after unrolling)
template <int SPT, int NB>
__device__ __forceinline__ double chain_pow(const double (&X)[NB][SPT][3], int nsrc, unsigned p, unsigned sb, int s, int j)
{
    double xv = X[0][s][j];
    if (NB >= 2 && nsrc >= 2 && sb == 1u) xv = X[NB >= 2 ? 1 : 0][s][j];
    if (NB >= 3 && nsrc >= 3 && sb == 2u) xv = X[NB >= 3 ? 2 : 0][s][j];
    return p == 1u ? xv : xv * xv;
}

// one term at the three collocation points
template <int SPT, int NB, bool CTRL, class CtrlLd>
__device__ __forceinline__ void chain_term(unsigned n_off, unsigned meta, int e, const char* __restrict__ base, CtrlLd& ctrl_ld,
                                           const double (&X)[NB][SPT][3], const int nsrc, double scale, double (&out)[SPT][3])
{
    if (!(meta & CT_PRESENT)) {
#pragma unroll
        for (int s = 0; s < SPT; ++s) out[s][0] = out[s][1] = out[s][2] = 0.0;
        return;
    }
    double nv[SPT];
    chain_narrow<SPT, CTRL>(n_off, meta, e, base, ctrl_ld, scale, nv);
    const unsigned p = (meta >> 8) & 3u, sb = (meta >> 12) & 3u;
    if (nsrc == 0 || p == 0u) {
#pragma unroll
        for (int s = 0; s < SPT; ++s) out[s][0] = out[s][1] = out[s][2] = nv[s];
        return;
    }
#pragma unroll
    for (int s = 0; s < SPT; ++s)
#pragma unroll
        for (int j = 0; j < 3; ++j) out[s][j] = nv[s] * chain_pow<SPT, NB>(X, nsrc, p, sb, s, j);
}

// NB = blocks the instantiation holds in registers (2 or 3); SB / SQ = the chain's compile-time shape (0: decoded at run time).
// Copies of the Newton iteration per loop trip: with two, the loop-carried iterates need no register moves at the back
// edge (measured on C3: 1 -> 16.60 ms, 2 -> 16.26, 3 -> 16.28, 4 -> 16.36).
#ifndef PYDSB_NEWTON_UNROLL
#define PYDSB_NEWTON_UNROLL 2
#endif
constexpr int NEWTON_UNROLL = PYDSB_NEWTON_UNROLL;

template <int SPT, int NB, unsigned long long SB, unsigned SQ, class CtrlLd>
__device__ __forceinline__ void poly_chain(const ExecImage& img, char* __restrict__ base, const char* __restrict__ cz,
                                           const double h, CtrlLd ctrl_ld, bool (&failed)[SPT],
                                           unsigned long long (&my_acc)[SPT], unsigned& cnt_warp_iters)
{
    typedef Vec<SPT> VecT;
    typedef ChainShape<SB, SQ> SH;
    constexpr unsigned USP = US * SPT;
    constexpr bool CTRL = SH::ctrl;
    const int nb = SH::nb(), nq = SH::nq();
    const int max_it = (int)c_chain.max_it;

    double X[NB][SPT][3];                                  // collocation points of every block (point 2 = element end)
    double Q[CHAIN_MAXQ][SPT];
    unsigned nit[NB][SPT];                                 // Newton iterations of every block over all elements
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        VecT v;
#pragma unroll
        for (int s = 0; s < SPT; ++s) v.v[s] = 0.0;
        if (b < nb) v = *reinterpret_cast<const VecT*>((c_chain.blk[b].x0_const ? cz : base) + c_chain.blk[b].x0_off);
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            X[b][s][0] = X[b][s][1] = X[b][s][2] = v.v[s];
            nit[b][s] = 0u;
        }
    }
#pragma unroll
    for (int q = 0; q < CHAIN_MAXQ; ++q) {
        VecT v;
#pragma unroll
        for (int s = 0; s < SPT; ++s) v.v[s] = 0.0;
        if (q < nq) v = *reinterpret_cast<const VecT*>((c_chain.quad[q].q0_const ? cz : base) + c_chain.quad[q].q0_off);
#pragma unroll
        for (int s = 0; s < SPT; ++s) Q[q][s] = v.v[s];
    }

    // ---- LINCONST blocks: (A - h c1 I)^-1 once per scenario, into units that are dead during the loop ----
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (b >= nb || SH::kind(b) != CHAIN_LINCONST) continue;
        VecT c1;
#pragma unroll
        for (int s = 0; s < SPT; ++s) c1.v[s] = 0.0;
        if (SH::meta(b, 1) & CT_PRESENT) c1 = LDV(c_chain.blk[b].c[1].n_off);   // narrow, control-free (assembler)
        VecT mi[9];
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            const double lam = (SH::meta(b, 1) & CT_NEG) ? -h * c1.v[s] : h * c1.v[s];
            double M[3][3];
#pragma unroll
            for (int jj = 0; jj < 3; ++jj)
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) M[jj][kk] = (jj == kk) ? CH_A(jj, kk) - lam : CH_A(jj, kk);
            // adjugate / determinant: no dependent chain of pivots (the matrix is the well-conditioned collocation matrix
            // shifted by a decay rate)
            double C[3][3];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                    C[i][j] = fma(M[i1][j1], M[i2][j2], -(M[i1][j2] * M[i2][j1]));      // cofactor (cyclic form: sign included)
                }
            const double det = fma(M[0][0], C[0][0], fma(M[0][1], C[0][1], M[0][2] * C[0][2]));
            const double rd = fast_rcp(det);
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) mi[3 * i + j].v[s] = C[j][i] * rd;                // inverse = adjugate^T / det
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) STV(c_chain.blk[b].minv_off + i * USP, mi[i]);
    }

    // ---- the finite elements ----
#pragma unroll 1
    for (int e = 0; e < img.nfe; ++e) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (b >= nb) continue;
            const ChainBlockD& bd = c_chain.blk[b];
            const unsigned kind = SH::kind(b);
            // h c0 at the three points, folded with the x(element start) column:  cb_j = h c0_j - A0_j x_s
            double cb[SPT][3], hc1[SPT][3], hc2[SPT];
            chain_term<SPT, NB, CTRL>(bd.c[0].n_off, SH::meta(b, 0), e, base, ctrl_ld, X, b, h, cb);
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                const double xs = X[b][s][2];
#pragma unroll
                for (int j = 0; j < 3; ++j) cb[s][j] = fma(-CH_A0(j), xs, cb[s][j]);
            }
            if (kind == CHAIN_LINCONST) {
                // x = (A - h c1 I)^-1 cb
                double xn[SPT][3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const VecT m0 = LDV(bd.minv_off + (3 * j + 0) * USP), m1 = LDV(bd.minv_off + (3 * j + 1) * USP),
                               m2 = LDV(bd.minv_off + (3 * j + 2) * USP);
#pragma unroll
                    for (int s = 0; s < SPT; ++s) {
                        double a = m0.v[s] * cb[s][0];
                        a = fma(m1.v[s], cb[s][1], a);
                        xn[s][j] = fma(m2.v[s], cb[s][2], a);
                    }
                }
#pragma unroll
                for (int s = 0; s < SPT; ++s) { X[b][s][0] = xn[s][0]; X[b][s][1] = xn[s][1]; X[b][s][2] = xn[s][2]; }
                continue;
            }
            chain_term<SPT, NB, CTRL>(bd.c[1].n_off, SH::meta(b, 1), e, base, ctrl_ld, X, b, h, hc1);
            {
                double z3[SPT][3];                          // c2: narrow (assembler)
                chain_term<SPT, NB, CTRL>(bd.c[2].n_off, SH::meta(b, 2), e, base, ctrl_ld, X, 0, h, z3);
#pragma unroll
                for (int s = 0; s < SPT; ++s) hc2[s] = z3[s][0];
            }
            const bool linear = kind == CHAIN_LINEAR;
            bool conv[SPT];
            double s_prev[SPT], lim[SPT], dg[SPT][3];
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                // stop test against |x(element start)|: one limit per element instead of |x|^2 of every iterate
                const double xs = X[b][s][2];
                lim[s] = fma(img.rtol2 * 3.0, xs * xs, img.rtol2);
                X[b][s][0] = X[b][s][1] = xs;               // start: x_j = x(element start)
                conv[s] = false;
                s_prev[s] = 0.0;
                // the diagonal of the linear part joins the Horner form:  h f(x_j) - A_jj x_j = (dg_j + hc2 x_j) x_j + h c0_j
#pragma unroll
                for (int j = 0; j < 3; ++j) dg[s][j] = hc1[s][j] - CH_A(j, j);
            }
            double lim_step[SPT];                                     // |dx|^2 below which 2^-16 |dx| is inside the tolerance
#pragma unroll
            for (int s = 0; s < SPT; ++s) lim_step[s] = PYDSB_STEP_RCP == 2 ? lim[s] * 0x1p32 : DBL_MAX;
#pragma unroll NEWTON_UNROLL
            for (int it = 0;;) {
                bool all = true;
#pragma unroll
                for (int s = 0; s < SPT; ++s) {
                    nit[b][s] += conv[s] ? 0u : 1u;
                    // N = h J - A (the negated Newton matrix):  N dx = r,  x -= dx
                    double N[3][3], r[3];
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj) {
                        const double x = X[b][s][jj];
                        const double t = fma(hc2[s], x, dg[s][jj]);
                        double rr = fma(t, x, cb[s][jj]);                 // h f(x_j) - A0_j x_s - A_jj x_j
#pragma unroll
                        for (int kk = 0; kk < 3; ++kk) {
                            if (kk != jj) rr = fma(-CH_A(jj, kk), X[b][s][kk], rr);
                            N[jj][kk] = -CH_A(jj, kk);
                        }
                        r[jj] = rr;
                        N[jj][jj] = fma(hc2[s], x, t);                   // h (2 c2 x + c1) - A_jj
                    }
                    // a Newton step may be inexact (pivot reciprocals: the bare 20-bit seed -- three dependent FMAs less per
                    // pivot in the one serial chain of the iteration); a linear block is solved by this one step: exact pivots
                    if (linear) lu_solve_inplace<3, false>(N, r);
                    else lu_solve_inplace<3, true>(N, r);
                    double sq = 0.0;
#pragma unroll
                    for (int kk = 0; kk < 3; ++kk) {
                        X[b][s][kk] -= r[kk];
                        sq = fma(r[kk], r[kk], sq);
                    }
                    // stop: the step is below the tolerance, or the NEXT one will be -- by the quadratic prediction
                    // |dx|^2 / |dx_prev| and by what the inexact pivots leave behind, 2^-16 |dx| (seed error x conditioning)
                    const bool ok = linear ? (sq <= DBL_MAX)
                                           : ((sq <= lim[s]) | ((sq * sq <= lim[s] * s_prev[s]) & (sq <= 0.01 * s_prev[s]) &
                                                                (sq <= lim_step[s])));
                    s_prev[s] = sq;
                    conv[s] = conv[s] | ok;
                    all = all & conv[s];
                }
                ++cnt_warp_iters;
                ++it;
                if (linear || __all_sync(0xffffffffu, all) || it >= max_it) break;
            }
#pragma unroll
            for (int s = 0; s < SPT; ++s) failed[s] = failed[s] | !conv[s];
        }
        // quadratures:  q += h sum_j bw_j fq(x_j)  at the converged points (sum_j bw_j = 1 for an element-constant integrand)
#pragma unroll
        for (int q = 0; q < CHAIN_MAXQ; ++q) {
            if (q >= nq) continue;
            const unsigned qm = SH::qmeta(q);
            double nv[SPT];
            chain_narrow<SPT, CTRL>(c_chain.quad[q].n_off, qm, e, base, ctrl_ld, h, nv);
            const unsigned p = (qm >> 8) & 3u, sb = (qm >> 12) & 3u;
            if (p == 0u) {
#pragma unroll
                for (int s = 0; s < SPT; ++s) Q[q][s] += nv[s];
            } else {
#pragma unroll
                for (int s = 0; s < SPT; ++s) {
                    double acc = img.bw[0] * chain_pow<SPT, NB>(X, NB, p, sb, s, 0);
                    acc = fma(img.bw[1], chain_pow<SPT, NB>(X, NB, p, sb, s, 1), acc);
                    acc = fma(img.bw[2], chain_pow<SPT, NB>(X, NB, p, sb, s, 2), acc);
                    Q[q][s] = fma(nv[s], acc, Q[q][s]);
                }
            }
        }
    }

    // ---- converged end values to the slots the g tape reads; counters ----
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (b >= nb) continue;
        const unsigned kind = SH::kind(b);
        VecT o;
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            o.v[s] = X[b][s][2];
            // a LINCONST block has no stop test: a non-finite value reaches the end (the recurrence is linear in x)
            failed[s] = failed[s] | !(fabs(X[b][s][2]) <= DBL_MAX);
            if (kind == CHAIN_LINCONST)
                my_acc[s] += ((unsigned long long)img.nfe << 40) | (unsigned long long)(CHAIN_INV_FLOPS + img.nfe * CHAIN_LINCONST_FLOPS);
            else
                my_acc[s] += ((unsigned long long)nit[b][s] << 40) | (unsigned long long)(nit[b][s] * c_chain.blk[b].flops);
        }
        STV(c_chain.blk[b].xe_off, o);
    }
#pragma unroll
    for (int q = 0; q < CHAIN_MAXQ; ++q) {
        if (q >= nq) continue;
        VecT v;
#pragma unroll
        for (int s = 0; s < SPT; ++s) v.v[s] = Q[q][s];
        STV(c_chain.quad[q].q_off, v);
    }
}

#undef CH_A
#undef CH_A0

}  // namespace pydsb
