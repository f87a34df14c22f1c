// rsens_kernel: state solve + forward control sensitivities of a polynomial chain, in registers.
//
// The evaluator of the batched control optimisation (recourse_kernel.cuh) for models whose element loop assembles to a
// fused chain (poly_chain.cuh): same inputs, same outputs and the same optimiser around it as sens_kernel, which
// interprets the coupled-block sensitivity program (a point-tape pass per Newton iteration, the sensitivity block in
// shared memory: 19 % FP64-pipe utilisation, profiles/r01_v23_sens_kernel_ncu_summary.txt).  Here, per scenario:
//
//   scalar ops (parameter tape)                                              as in the chain instantiation of feas_kernel
//   per finite element, per block: coefficients, Newton / one matrix-vector product      (poly_chain.cuh)
//                                  + the LU factors of the Newton matrix at the converged points
//   per finite element, per control that has acted so far, per block:
//        N dX = A0 S_prev - d(h c0) - d(h c1) x        with the block's factors (a constant-decay block: its inverse)
//        S = dX at the element end; the quadrature sensitivities alongside
//   scalar ops (g tape with dg/dxe, dg/dqe), c_k = g_k / M_k, grad c_k = (dg/dxe S + dg/dqe Sq) / M_k, recourse_emit.
//
// Every state, sensitivity (ns x nz + nq x nz) and factor lives in registers; only the narrow coefficients and the
// inverse of a constant-decay block are read from the shared-memory register file.  What replaces what in the reference:
// the NLP solve pyds/solver.py:50-55 hands to GAMS/CONOPT (derivatives by the solver's own differentiation of the GAMS
// model).
//
// Supported shape (the assembler decides, pydsb_api.cu: otherwise sens_kernel runs): <= RC_NB blocks, <= RC_NQ
// quadratures, <= RC_NZ control values, a control enters a coefficient or integrand as (narrow) x (control), c2 narrow.
#pragma once
#include "poly_chain.cuh"
#include "recourse_kernel.cuh"

namespace pydsb {

constexpr int RC_NZ = 8;                   // control values whose sensitivities are carried in registers
constexpr int RC_NQ = 4;
constexpr int RC_MAXFE = 32;
constexpr unsigned RC_REF_NONE = 0xFFFFFFFFu, RC_REF_CONST = 0x80000000u;

// Program 2 of a model blob (the feasibility program with dg/dxe, dg/dqe in its g tape), assembled for one scenario per
// thread, and what the sensitivity propagation needs beside the chain descriptor
constexpr int MAX_PROG_R = 512;
__constant__ uint4 c_prog_r[2 * MAX_PROG_R];
struct RChainDesc {
    ChainDesc ch;
    unsigned gx_off[8][CHAIN_MAXB];        // dg_k / d(end value of block b): slot offset, | RC_REF_CONST: constant pool, NONE: 0
    unsigned gq_off[8][RC_NQ];
    unsigned char cmap[RC_MAXFE][4];       // z index of control function k on finite element e
    unsigned live[RC_MAXFE];               // bit c: control value c has acted on an element <= e
    unsigned zfix_mask;                    // bit c: control value c is fixed (no sensitivity wanted) ...
    unsigned nce, pad0, pad1;
    unsigned zfix_ref[RC_NZ];              // ... and comes from the model (a parameter expression): slot / constant reference
};
__constant__ RChainDesc c_rchain;

// compile-time shape of the chain (cf. ChainShape in poly_chain.cuh; SB = 0: decoded at run time)
template <unsigned long long SB, unsigned SQ>
struct RChainShape {
    static constexpr bool generic = SB == 0ull;
    static constexpr int NQC = generic ? RC_NQ : (int)(SQ & 7u);       // quadratures whose sensitivities are carried
    static __device__ __forceinline__ int nb() { return generic ? (int)c_rchain.ch.nb : (int)(SB & 3ull); }
    static __device__ __forceinline__ int nq() { return generic ? (int)c_rchain.ch.nq : (int)(SQ & 7u); }
    static __device__ __forceinline__ unsigned kind(int b)
    {
        return generic ? c_rchain.ch.blk[b].kind : (unsigned)((SB >> (3 + 19 * b)) & 3ull);
    }
    static __device__ __forceinline__ unsigned meta(int b, int k)
    {
        const unsigned m = c_rchain.ch.blk[b].c[k].meta;
        if (generic) return m;
        unsigned t = (unsigned)((SB >> (3 + 19 * b + 2 + 7 * k)) & (k == 2 ? 7ull : 127ull));
        if (k == 2) t = (t & 3u) | ((t >> 2) << 6);
        return ChainShape<SB, SQ>::unpack(t) | (m & (CT_HAS_CTRL | 0xF0u));
    }
    static __device__ __forceinline__ unsigned qmeta(int q)
    {
        const unsigned m = c_rchain.ch.quad[q].meta;
        if (generic) return m;
        return ChainShape<SB, SQ>::unpack((SQ >> (3 + 7 * q)) & 127u) | (m & (CT_HAS_CTRL | 0xF0u));
    }
};

#define RC_A(jj, kk) img.adot[4 * ((kk) + 1) + (jj) + 1]
#define RC_A0(jj) img.adot[(jj) + 1]
#define RC_LD(off) (*reinterpret_cast<const double*>(base + (off)))
#define RC_LD_W(b_, off) (*reinterpret_cast<double*>((b_) + (off)))

// value of a narrow reference of the descriptor: slot of this thread, constant-pool entry, or absent (0)
__device__ __forceinline__ double rc_ref(const char* __restrict__ base, const char* __restrict__ cz, unsigned ref)
{
    if (ref == RC_REF_NONE) return 0.0;
    return (ref & RC_REF_CONST) ? *reinterpret_cast<const double*>(cz + (ref & ~RC_REF_CONST)) : RC_LD(ref);
}

// scalar section of a chain program, one scenario per thread (cf. the chain instantiation of feas_kernel)
__device__ __forceinline__ void rc_scalar_run(char* __restrict__ base, const char* __restrict__ cz, int pc, const int pc_end)
{
    for (; pc < pc_end; pc += 2) {
        const uint4 w0 = c_prog_r[pc], w1 = c_prog_r[pc + 1];
        const unsigned fl = w1.y, op = w0.x;
        const double a0 = *reinterpret_cast<const double*>(((fl & 1u) ? cz : base) + w0.z);
        const double b = *reinterpret_cast<const double*>(((fl & 2u) ? cz : base) + w0.w);
        double r;
        if (op == X_FMAG) {
            const double c0 = *reinterpret_cast<const double*>(((fl & 4u) ? cz : base) + w1.x);
            const int sa = (int)((fl & 8u) << 28), sc = (int)((fl & 16u) << 27);
            const double a = __hiloint2double(__double2hiint(a0) ^ sa, __double2loint(a0));
            const double c = __hiloint2double(__double2hiint(c0) ^ sc, __double2loint(c0));
            r = fma(a, b, c);
        } else if (op == X_COLD1 && w1.z == OP_EXP) {
            r = exp(a0);
        } else {
            switch (op) {
            case X_DIV1: { const double rb = fast_rcp(b), q = a0 * rb; r = fma(fma(-b, q, a0), rb, q); break; }
            case X_RCP1: { const double ra = fast_rcp(a0); r = fma(fma(-a0, ra, 1.0), ra, ra); break; }
            case X_ABS1: r = fabs(a0); break;
            case X_MIN1: r = fmin(a0, b); break;
            case X_MAX1: r = fmax(a0, b); break;
            default: r = cold_op(w1.z, a0, b); break;
            }
        }
        *reinterpret_cast<double*>(base + w0.y) = r;
    }
}

// one coefficient / integrand term on finite element e:  +-scale * narrow [* z] * (state of block src)^p
struct RcTerm {
    double nvz;        // the narrow factor including the control value (what multiplies the power)
    double nvn;        // the narrow factor without the control value: d(term) / d(control) = nvn * power;  0 without control
    int zi;            // z index the control maps to on this element (-1: no control)
    unsigned p, src;
};
__device__ __forceinline__ RcTerm rc_term(const char* __restrict__ base, const char* __restrict__ cz, unsigned n_off, unsigned meta,
                                          int e, const double* __restrict__ zs, double scale)
{
    RcTerm t;
    t.nvz = t.nvn = 0.0; t.zi = -1; t.p = (meta >> 8) & 3u; t.src = (meta >> 12) & 3u;
    if (!(meta & CT_PRESENT)) return t;
    double nv = (meta & CT_NEG) ? -scale : scale;
    if (meta & CT_HAS_N) nv *= RC_LD(n_off);
    t.nvz = nv;
    if (meta & CT_HAS_CTRL) {
        t.zi = (int)c_rchain.cmap[e][(meta >> 4) & 0xFu];
        t.nvn = nv;
        t.nvz = nv * (((c_rchain.zfix_mask >> t.zi) & 1u) ? rc_ref(base, cz, c_rchain.zfix_ref[t.zi]) : zs[(size_t)t.zi * BLOCK]);
    }
    return t;
}

template <int NB>
__device__ __forceinline__ double rc_pow(const double (&X)[NB][3], unsigned p, unsigned sb, int j)
{
    double xv = X[0][j];
    if (NB >= 2 && sb == 1u) xv = X[NB >= 2 ? 1 : 0][j];
    if (NB >= 3 && sb == 2u) xv = X[NB >= 3 ? 2 : 0][j];
    return p == 0u ? 1.0 : (p == 1u ? xv : xv * xv);
}
// d(power)/d(state) * dX:  p x^(p-1) dx
template <int NB>
__device__ __forceinline__ double rc_dpow(const double (&X)[NB][3], const double (&dX)[NB][3], unsigned p, unsigned sb, int j)
{
    if (p == 0u) return 0.0;
    double xv = X[0][j], dv = dX[0][j];
    if (NB >= 2 && sb == 1u) { xv = X[NB >= 2 ? 1 : 0][j]; dv = dX[NB >= 2 ? 1 : 0][j]; }
    if (NB >= 3 && sb == 2u) { xv = X[NB >= 3 ? 2 : 0][j]; dv = dX[NB >= 3 ? 2 : 0][j]; }
    return p == 1u ? dv : 2.0 * xv * dv;
}

// la == nullptr: one evaluation (the host loop of pydsb_recourse_solve launches lm_update_kernel next).
// la != nullptr (rfused_kernel): the WHOLE optimisation of a shared-control group whose theta set fits one tile -- evaluate,
// reduce the statistics over the CTA, thread 0 takes the Levenberg-Marquardt decision (lm_update_group_regs, the same
// code the separate kernel runs), repeat until the group stops: no launch, no host round trip per iteration.
template <int NB, unsigned long long SB, unsigned SQ>
__device__ __forceinline__ void rsens_body(const ExecImage& img, const RecourseArgs& ra, const LmArgs* la)
{
    typedef RChainShape<SB, SQ> SH;
    constexpr int NQ = SH::NQC;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const ChainDesc& cd = c_rchain.ch;

    double* cpool = reinterpret_cast<double*>(smem + img.off_const) + 3;        // chain layout: 1.0, -0.0, 0.0, then the pool
    double* wsum = reinterpret_cast<double*>(smem + img.off_slots);
    double* S = reinterpret_cast<double*>(smem + img.off_slots) + tid;
    for (int i = tid; i < img.n_const; i += BLOCK) cpool[i] = img.consts[i];
    if (tid == 0) { cpool[-3] = 1.0; cpool[-2] = -0.0; cpool[-1] = 0.0; }
    __syncthreads();
    char* const base = reinterpret_cast<char*>(S);
    const char* const cz = reinterpret_cast<const char*>(cpool - 3);
    const int nz = img.nz, ng = img.ng;
    const int nb = SH::nb(), nq = SH::nq();
    double* const Dc = S + (size_t)ra.sens_slot * BLOCK;          // Dc[(k * nz + c) * BLOCK], one spare row after the ng rows
    const double h = 1.0 / (double)img.nfe;
    const int NV = recourse_nv(nz);

  for (int fused_it = 0;; ++fused_it) {
    // ---- which scenario (as sens_kernel) ----
    const long long t = blockIdx.x;
    const bool per_scen = ra.per_scenario == 1;
    long long i_d, j, grp;
    bool valid;
    if (per_scen) {
        const long long n_run = ra.n_active ? (long long)*ra.n_active : ra.n_d * ra.n_t;
        const long long k0 = t * (long long)BLOCK;
        if (k0 >= n_run) return;
        valid = k0 + tid < n_run;
        const long long k = valid ? k0 + tid : n_run - 1;
        grp = ra.active ? ra.active[k] : k;
        i_d = grp / ra.n_t;
        j = grp % ra.n_t;
    } else if (ra.per_scenario == 2) {
        const long long k0 = t * (long long)BLOCK;
        const long long rem = ra.n_d * ra.n_t - k0;
        const int rows = rem < BLOCK ? (int)rem : BLOCK;
        valid = tid < rows;
        const long long k = k0 + (valid ? tid : rows - 1);
        i_d = k / ra.n_t;
        j = k % ra.n_t;
        grp = 0;
        if (ra.status[0] != 0) return;
    } else {
        i_d = t / ra.n_tiles;
        const long long j0 = (t % ra.n_tiles) * (long long)BLOCK;
        const long long rem = ra.n_t - j0;
        const int rows = rem < BLOCK ? (int)rem : BLOCK;
        valid = tid < rows;
        j = j0 + (valid ? tid : rows - 1);
        grp = i_d;
        if (__ldcg(ra.status + grp) != 0) return;
    }
    const bool active = valid && __ldcg(ra.status + grp) == 0;
    if (!__syncthreads_or(active)) return;

    for (int c = 0; c < img.p_dim; ++c) RC_LD_W(base, cd.par_off[img.d_dim + c]) = __ldg(ra.theta + j * img.p_dim + c);
    for (int c = 0; c < img.d_dim; ++c) RC_LD_W(base, cd.par_off[c]) = __ldg(ra.d + i_d * img.d_dim + c);
    // the group's control values, staged once per evaluation in this thread's column of the (still unused) Dc rows: one
    // round of L2 loads in flight together instead of an L2 latency per term and element (z_try is rewritten between the
    // evaluations of the fused loop by another thread: read past L1)
    double* const zs = Dc;
    {
        const double* const zg = ra.z_try + grp * nz;
        double zv[RC_NZ];
#pragma unroll
        for (int c = 0; c < RC_NZ; ++c) zv[c] = c < nz ? __ldcg(zg + c) : 0.0;
#pragma unroll
        for (int c = 0; c < RC_NZ; ++c)
            if (c < nz) zs[(size_t)c * BLOCK] = zv[c];
    }

    rc_scalar_run(base, cz, 0, (int)cd.pre_end);

    // ---- the chain: states X, quadratures Q, their sensitivities Sx, Sq with respect to the nz control values ----
    double X[NB][3], Sx[NB][RC_NZ], Q[NQ], Sq[NQ][RC_NZ];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const double v = b < nb ? *reinterpret_cast<const double*>((cd.blk[b].x0_const ? cz : base) + cd.blk[b].x0_off) : 0.0;
        X[b][0] = X[b][1] = X[b][2] = v;
#pragma unroll
        for (int c = 0; c < RC_NZ; ++c) Sx[b][c] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        Q[q] = q < nq ? *reinterpret_cast<const double*>((cd.quad[q].q0_const ? cz : base) + cd.quad[q].q0_off) : 0.0;
#pragma unroll
        for (int c = 0; c < RC_NZ; ++c) Sq[q][c] = 0.0;
    }
    // constant-decay blocks: (A - h c1 I)^-1 once per scenario (adjugate / determinant), 9 values in dead units
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (b >= nb || SH::kind(b) != CHAIN_LINCONST) continue;
        const unsigned m1 = SH::meta(b, 1);
        double lam = (m1 & CT_PRESENT) ? h * RC_LD(cd.blk[b].c[1].n_off) : 0.0;
        if (m1 & CT_NEG) lam = -lam;
        double M[3][3], C[3][3];
#pragma unroll
        for (int jj = 0; jj < 3; ++jj)
#pragma unroll
            for (int kk = 0; kk < 3; ++kk) M[jj][kk] = (jj == kk) ? RC_A(jj, kk) - lam : RC_A(jj, kk);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int jx = 0; jx < 3; ++jx) {
                const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (jx + 1) % 3, j2 = (jx + 2) % 3;
                C[i][jx] = fma(M[i1][j1], M[i2][j2], -(M[i1][j2] * M[i2][j1]));
            }
        const double rd = fast_rcp(fma(M[0][0], C[0][0], fma(M[0][1], C[0][1], M[0][2] * C[0][2])));
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int jx = 0; jx < 3; ++jx) RC_LD_W(base, cd.blk[b].minv_off + (3 * i + jx) * US) = C[jx][i] * rd;
    }

    bool failed = false;
    unsigned my_iters = 0;
#pragma unroll 1
    for (int e = 0; e < img.nfe; ++e) {
        // per block: the two coefficient terms that can carry a control or an earlier state, the LU factors of the Newton
        // matrix at the converged points
        RcTerm t0[NB], t1[NB];
        double fac[NB][3][3], finv[NB][3];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            if (b >= nb) continue;
            const ChainBlockD& bd = cd.blk[b];
            const unsigned kind = SH::kind(b);
            t0[b] = rc_term(base, cz, bd.c[0].n_off, SH::meta(b, 0), e, zs, h);
            t1[b] = rc_term(base, cz, bd.c[1].n_off, SH::meta(b, 1), e, zs, h);
            const RcTerm t2 = rc_term(base, cz, bd.c[2].n_off, SH::meta(b, 2), e, zs, h);     // narrow, control free (assembler)
            const double hc2 = t2.nvz;
            const double xs = X[b][2];
            double cb[3], dg[3];
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
                cb[jj] = fma(-RC_A0(jj), xs, t0[b].nvz * rc_pow<NB>(X, b > 0 ? t0[b].p : 0u, t0[b].src, jj));
                dg[jj] = t1[b].nvz * rc_pow<NB>(X, b > 0 ? t1[b].p : 0u, t1[b].src, jj) - RC_A(jj, jj);
            }
            if (kind == CHAIN_LINCONST) {
                double xn[3];
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    double a = RC_LD(bd.minv_off + (3 * jj + 0) * US) * cb[0];
                    a = fma(RC_LD(bd.minv_off + (3 * jj + 1) * US), cb[1], a);
                    xn[jj] = fma(RC_LD(bd.minv_off + (3 * jj + 2) * US), cb[2], a);
                }
                X[b][0] = xn[0]; X[b][1] = xn[1]; X[b][2] = xn[2];
                continue;
            }
            const bool linear = kind == CHAIN_LINEAR;
            const double lim = fma(img.rtol2 * 3.0, xs * xs, img.rtol2);
            const double lim_step = PYDSB_STEP_RCP == 2 ? lim * 0x1p32 : DBL_MAX;
            X[b][0] = X[b][1] = xs;
            bool conv = !active;
            double s_prev = 0.0;
#pragma unroll NEWTON_UNROLL
            for (int it = 0;;) {
                double N[3][3], r[3];
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    const double x = X[b][jj];
                    const double tt = fma(hc2, x, dg[jj]);
                    double rr = fma(tt, x, cb[jj]);
#pragma unroll
                    for (int kk = 0; kk < 3; ++kk) {
                        if (kk != jj) rr = fma(-RC_A(jj, kk), X[b][kk], rr);
                        N[jj][kk] = -RC_A(jj, kk);
                    }
                    r[jj] = rr;
                    N[jj][jj] = fma(hc2, x, tt);
                }
                // inexact Newton step / exact single step of a linear block, and the stop test, as in poly_chain.cuh
                if (linear) lu_solve_inplace<3, false>(N, r);
                else lu_solve_inplace<3, true>(N, r);
                double sq = 0.0;
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) { X[b][kk] -= r[kk]; sq = fma(r[kk], r[kk], sq); }
                if (!conv) ++my_iters;
                const bool bad = !(sq <= DBL_MAX);
                const bool ok = linear ? true : ((sq <= lim) | ((sq * sq <= lim * s_prev) & (sq <= 0.01 * s_prev) & (sq <= lim_step)));
                s_prev = sq;
                if (bad && active) failed = true;
                conv = conv | ok | bad;
                ++it;
                if (linear || __all_sync(0xffffffffu, conv) || it >= img.max_it) break;
            }
            if (!conv) failed = true;
            // factors of N at the converged points: every sensitivity system of this element reuses them
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) {
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) fac[b][jj][kk] = -RC_A(jj, kk);
                fac[b][jj][jj] = fma(2.0 * hc2, X[b][jj], dg[jj]);
            }
            lu_factor<3>(fac[b], finv[b]);
        }
        // quadrature terms of this element, then the quadratures themselves
        RcTerm tq[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            if (q >= nq) continue;
            tq[q] = rc_term(base, cz, cd.quad[q].n_off, SH::qmeta(q), e, zs, h);
            double acc = img.bw[0] * rc_pow<NB>(X, tq[q].p, tq[q].src, 0);
            acc = fma(img.bw[1], rc_pow<NB>(X, tq[q].p, tq[q].src, 1), acc);
            acc = fma(img.bw[2], rc_pow<NB>(X, tq[q].p, tq[q].src, 2), acc);
            Q[q] = fma(tq[q].nvz, acc, Q[q]);
        }

        // ---- forward sensitivities with respect to every control value that has acted so far ----
        const unsigned live = c_rchain.live[e] & ~c_rchain.zfix_mask;
#pragma unroll
        for (int c = 0; c < RC_NZ; ++c) {
            if (!((live >> c) & 1u)) continue;
            double dX[NB][3];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                dX[b][0] = dX[b][1] = dX[b][2] = 0.0;
                if (b >= nb) continue;
                // d(h c0_j) + d(h c1_j) x_j :  the control of this element directly, earlier blocks through their points
                double dc[3];
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
                    double d0 = t0[b].nvz * (b > 0 ? rc_dpow<NB>(X, dX, t0[b].p, t0[b].src, jj) : 0.0);
                    double d1 = t1[b].nvz * (b > 0 ? rc_dpow<NB>(X, dX, t1[b].p, t1[b].src, jj) : 0.0);
                    if (t0[b].zi == c) d0 = fma(t0[b].nvn, rc_pow<NB>(X, b > 0 ? t0[b].p : 0u, t0[b].src, jj), d0);
                    if (t1[b].zi == c) d1 = fma(t1[b].nvn, rc_pow<NB>(X, b > 0 ? t1[b].p : 0u, t1[b].src, jj), d1);
                    dc[jj] = fma(d1, X[b][jj], d0);
                }
                if (SH::kind(b) == CHAIN_LINCONST) {
                    // x = Minv (h c0 - A0 x_s)  ->  dX = Minv (d(h c0) - A0 S_prev)
                    double rh[3];
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj) rh[jj] = fma(-RC_A0(jj), Sx[b][c], dc[jj]);
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj) {
                        double a = RC_LD(cd.blk[b].minv_off + (3 * jj + 0) * US) * rh[0];
                        a = fma(RC_LD(cd.blk[b].minv_off + (3 * jj + 1) * US), rh[1], a);
                        dX[b][jj] = fma(RC_LD(cd.blk[b].minv_off + (3 * jj + 2) * US), rh[2], a);
                    }
                } else {
                    // N dX = A0 S_prev - d(h f)
                    double rh[3];
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj) rh[jj] = fma(RC_A0(jj), Sx[b][c], -dc[jj]);
                    lu_apply<3>(fac[b], finv[b], rh);
                    dX[b][0] = rh[0]; dX[b][1] = rh[1]; dX[b][2] = rh[2];
                }
                Sx[b][c] = dX[b][2];
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                if (q >= nq) continue;
                double acc = img.bw[0] * rc_dpow<NB>(X, dX, tq[q].p, tq[q].src, 0);
                acc = fma(img.bw[1], rc_dpow<NB>(X, dX, tq[q].p, tq[q].src, 1), acc);
                acc = fma(img.bw[2], rc_dpow<NB>(X, dX, tq[q].p, tq[q].src, 2), acc);
                double s = fma(tq[q].nvz, acc, Sq[q][c]);
                if (tq[q].zi == c) {
                    double pw = img.bw[0] * rc_pow<NB>(X, tq[q].p, tq[q].src, 0);
                    pw = fma(img.bw[1], rc_pow<NB>(X, tq[q].p, tq[q].src, 1), pw);
                    pw = fma(img.bw[2], rc_pow<NB>(X, tq[q].p, tq[q].src, 2), pw);
                    s = fma(tq[q].nvn, pw, s);
                }
                Sq[q][c] = s;
            }
        }
    }

    // ---- end values -> g tape (with dg/dxe, dg/dqe) -> scaled constraints and their control gradients ----
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        if (b >= nb) continue;
        RC_LD_W(base, cd.blk[b].xe_off) = X[b][2];
        if (!(fabs(X[b][2]) <= DBL_MAX) && active) failed = true;
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q)
        if (q < nq) RC_LD_W(base, cd.quad[q].q_off) = Q[q];
    rc_scalar_run(base, cz, (int)cd.g_begin, (int)cd.g_end);
    double cmax = 0.0;
    double ck[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        ck[k] = 0.0;
        if (k < ng) {
            ck[k] = RC_LD(cd.g_off[k]) * img.bigm_inv[k];
            cmax = fmax(cmax, ck[k]);
        }
    }
    double gxv[8][NB], gqv[8][NQ];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int b = 0; b < NB; ++b) gxv[k][b] = (k < ng && b < nb) ? rc_ref(base, cz, c_rchain.gx_off[k][b]) : 0.0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) gqv[k][q] = (k < ng && q < nq) ? rc_ref(base, cz, c_rchain.gq_off[k][q]) : 0.0;
    }
#pragma unroll
    for (int c = 0; c < RC_NZ; ++c) {
        if (c >= nz) continue;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k >= ng) continue;
            double v = 0.0;
#pragma unroll
            for (int b = 0; b < NB; ++b) v = fma(gxv[k][b], Sx[b][c], v);
#pragma unroll
            for (int q = 0; q < NQ; ++q) v = fma(gqv[k][q], Sq[q][c], v);
            Dc[(size_t)(k * nz + c) * BLOCK] = ((c_rchain.zfix_mask >> c) & 1u) ? 0.0 : v * img.bigm_inv[k];
        }
    }
    const int scratch_rows = img.n_units < 32 ? img.n_units : 32;
    recourse_emit(ra, nz, ng, NV, active, failed, ck, cmax, Dc, wsum, scratch_rows, grp, j, t, my_iters);
    if (la == nullptr) return;
    // ---- fused: the group's update by one thread, then the next evaluation (or the end).  One smoothing stage is emitted
    // per call (ra.n_stages = 1); when the update moves on to the next stage at the same controls it asks for that stage's
    // statistics (want_stage): c_k and their gradients are still in place, only the smoothing changes ----
    int* const want_stage = reinterpret_cast<int*>(smem + 64);
    for (bool resumed = false;; resumed = true) {
        __threadfence_block();
        __syncthreads();
        if (tid == 0) {
            *want_stage = 0;
            lm_update_group_regs<RC_NZ>(*la, grp, ra.n_stages == 1 ? want_stage : nullptr, resumed);
        }
        __threadfence_block();
        __syncthreads();
        if (!*want_stage) break;
        recourse_emit(ra, nz, ng, NV, active, failed, ck, cmax, Dc, wsum, scratch_rows, grp, j, t, my_iters, false);
    }
    if (fused_it > la->max_iter) return;                     // (the update stops the group at its iteration cap)
  }
}

template <int NB, unsigned long long SB = 0ull, unsigned SQ = 0u>
#ifndef PYDSB_RSENS_CTAS
#define PYDSB_RSENS_CTAS 3
#endif
__global__ void __launch_bounds__(BLOCK, SB ? PYDSB_RSENS_CTAS : 2) rsens_kernel(const __grid_constant__ ExecImage img, const __grid_constant__ RecourseArgs ra)
{
    rsens_body<NB, SB, SQ>(img, ra, nullptr);
}

template <int NB, unsigned long long SB = 0ull, unsigned SQ = 0u>
__global__ void __launch_bounds__(BLOCK, 2) rfused_kernel(const __grid_constant__ ExecImage img, const __grid_constant__ RecourseArgs ra,
                                                         const __grid_constant__ LmArgs la)
{
    rsens_body<NB, SB, SQ>(img, ra, &la);
}

#undef RC_A
#undef RC_A0
#undef RC_LD
#undef RC_LD_W

}  // namespace pydsb
