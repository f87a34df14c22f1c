// Feasibility kernel: collocation state solve + CQAs + big-M indicator + P(feasible) partials.
//
// Work decomposition (B200: 148 SMs, persistent CTAs):
//   tile   = one design point d_i  x  SPT*BLOCK consecutive theta samples
//   CTA    = BLOCK threads, loops over tiles  t = blockIdx.x, +gridDim.x, ...
//   thread = SPT (1 or 2) (d_i, theta_j) scenarios for the whole pipeline; with SPT = 2 the two scenarios of
//            a thread sit side by side in the register file (one LDS.128 / STS.128 moves both) and share
//            every fetch, dispatch and address of the interpreter
//
// Per scenario (what the reference does inside one GAMS/CONOPT solve per design point,
// pyds/run.py:55-67 + pyds/solver.py:50-55, for fixed controls) the thread runs ONE program, assembled
// by pydsb_model_create (assembler.h) from the lowered tapes and kept pre-decoded in the constant bank:
//
//     pro ops (parameter-only values: Arrhenius factors, ...)            INIT (initial states)
//     element loop:  LDCTRL..., elem ops,
//                    per block of the (block-triangular) state graph:
//                        NB   start the block's Newton solve from the element's initial value
//                        point ops: f and df/dx of the block at the 3 Radau points, evaluated 3-wide
//                        NS   residual + Jacobian assembled in registers -> register-resident dense LU
//                             -> update -> stop test; loops back to the block's point ops
//                    quadrature ops, QACC (Radau weights)                 ELEM_END
//     XEND, g ops (CQAs), END
//
// then (hand-written epilogue) indicator = any(g_k > 0)   (utils.py:23-29, solver.py:71-80) and
// w_j * 1[feasible] reduced with warp shuffles; one partial per tile (deterministic finish in
// reduce_partials_kernel) -- DEUS's efp reduction.
//
// The interpreter: the program counter and every decoded field live in uniform registers (LDCU
// fetch), the dispatch is one dense switch, and every opcode is specialised on its operand pattern
// (N = narrow: one value per scenario, W = wide: one value per collocation point, 1 KiB apart), so a
// body is nothing but LDS [thread base + uniform offset (+ imm)], the DFMAs, and STS.
//
// theta tiles are staged by 1-D bulk TMA (cp.async.bulk + mbarrier), double buffered: the copy
// for tile k+1 is in flight while tile k is computed.
#pragma once
#include "xvm_ops.cuh"
#include "poly_chain.cuh"

namespace pydsb {

// ---------------------------------------------------------------------------------------------
// Newton solve of one block (NSB states, N = 3*NSB collocation unknowns) on one finite element.
// newton_begin: start from the end point of the previous element (the INIT op seeds point 2 with x(0)).
// newton_step:  residual and Jacobian assembled in registers from the block's point-op outputs, dense LU
//               without pivoting (register resident), update.
// ---------------------------------------------------------------------------------------------
template <int NSB, int MAXNS, int SPT>
__device__ __forceinline__ void newton_begin(const unsigned (&xo)[3], char* __restrict__ base, double (&xs)[SPT][MAXNS],
                                             double (&zv)[SPT][3][MAXNS])
{
    typedef Vec<SPT> VecT;
    constexpr unsigned USP = US * SPT;
#pragma unroll
    for (int n = 0; n < NSB; ++n) {
        const VecT v = LDVP(xo[n], 2);
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            xs[s][n] = v.v[s];
            zv[s][0][n] = zv[s][1][n] = zv[s][2][n] = 0.0;    // increments z_j = x_j - x(element start)
        }
        STVP(xo[n], 0, v);
        STVP(xo[n], 1, v);
    }
}

// word k of the operand table that follows an NS instruction's first uint4
template <int NW>
struct NsTab {
    uint4 q[NW];
    __device__ __forceinline__ unsigned operator[](int k) const
    {
        const uint4& v = q[k >> 2];
        return (k & 3) == 0 ? v.x : ((k & 3) == 1 ? v.y : ((k & 3) == 2 ? v.z : v.w));
    }
};

// One Newton step on the collocation equations in increment form (the rows of the differentiation matrix
// sum to zero, so  sum_k adot[k][j] x_k = sum_{k>=1} adot[k][j] z_k):
//     r_j = h f(x_j) - sum_k A_kj z_k,    (A^T (x) I - h blockdiag(df/dx)) dz = r,    z += dz,   x_j = x_s + z_j.
// MASKED: skip the structurally zero Jacobian entries (bit n*NSB+m of jmask); 1-state blocks always load J.
// Returns |dz|^2 and |x|^2 per scenario for the stop test.
template <int NSB, int MAXNS, bool MASKED, int SPT>
__device__ __forceinline__ void newton_step(const ExecImage& img, int at, unsigned jmask, char* __restrict__ base,
                                            double h, const double (&xs)[SPT][MAXNS], double (&zv)[SPT][3][MAXNS],
                                            double (&sq)[SPT], double (&sc)[SPT])
{
    typedef Vec<SPT> VecT;
    constexpr int N = 3 * NSB;
    constexpr int NW = ns_length(NSB) - 1;
    NsTab<NW> tab;
#pragma unroll
    for (int i = 0; i < NW; ++i) tab.q[i] = c_prog[at + 1 + i];
    // table: x_o[NSB] | f_o[NSB][3] | j_o[NSB*NSB][3]
    constexpr int F0 = NSB, J0 = NSB + 3 * NSB;
    // SPT = 2: one 128-bit load serves both scenarios, so the operands are fetched up front; SPT = 1 loads each
    // operand where the assembly consumes it (the 6x6 / 9x9 systems have no registers to spare)
    constexpr bool PRE = SPT > 1;
    VecT fv[PRE ? NSB : 1][3], jv[PRE ? NSB * NSB : 1][3];
    if constexpr (PRE) {
#pragma unroll
        for (int n = 0; n < NSB; ++n)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj) fv[n][jj] = LDV(tab[F0 + 3 * n + jj]);
#pragma unroll
        for (int i = 0; i < NSB * NSB; ++i)
#pragma unroll
            for (int jj = 0; jj < 3; ++jj)
                if (!MASKED || ((jmask >> i) & 1u)) jv[i][jj] = LDV(tab[J0 + 3 * i + jj]);
    }
    VecT xn[NSB][3];
#pragma unroll
    for (int s = 0; s < SPT; ++s) {
        double M[N][N], r[N];
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
#pragma unroll
            for (int n = 0; n < NSB; ++n) {
                double lin = img.adot[4 + jj + 1] * zv[s][0][n];
                lin = fma(img.adot[8 + jj + 1], zv[s][1][n], lin);
                lin = fma(img.adot[12 + jj + 1], zv[s][2][n], lin);
                const double fval = PRE ? fv[PRE ? n : 0][jj].v[s] : LDV(tab[F0 + 3 * n + jj]).v[s];
                r[jj * NSB + n] = fma(h, fval, -lin);
#pragma unroll
                for (int kk = 0; kk < 3; ++kk)
#pragma unroll
                    for (int m = 0; m < NSB; ++m) {
                        const double a = (n == m) ? img.adot[4 * (kk + 1) + jj + 1] : 0.0;
                        M[jj * NSB + n][kk * NSB + m] = a;
                        if (kk == jj && (!MASKED || ((jmask >> (n * NSB + m)) & 1u)))
                            M[jj * NSB + n][kk * NSB + m] =
                                fma(-h, PRE ? jv[PRE ? n * NSB + m : 0][jj].v[s] : LDV(tab[J0 + 3 * (n * NSB + m) + jj]).v[s], a);
                    }
            }
        }

        lu_solve_inplace<N>(M, r);

        sq[s] = 0.0;
        sc[s] = 0.0;
#pragma unroll
        for (int kk = 0; kk < 3; ++kk)
#pragma unroll
            for (int n = 0; n < NSB; ++n) {
                const double dz = r[kk * NSB + n];
                const double zn = zv[s][kk][n] + dz;
                const double x = xs[s][n] + zn;
                zv[s][kk][n] = zn;
                xn[n][kk].v[s] = x;
                sq[s] = fma(dz, dz, sq[s]);
                sc[s] = fma(x, x, sc[s]);
            }
    }
    constexpr unsigned USP = US * SPT;
#pragma unroll
    for (int n = 0; n < NSB; ++n)
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) STVP(tab[n], kk, xn[n][kk]);
}

// MAXNS = largest block of coupled states (1..3); SPT = scenarios per thread (2 only with MAXNS = 1: the
// Newton matrices of two scenarios have to fit the register file).
// TRAJ (pydsb_state_solve, cold): the same program, but instead of g the state trajectories are written --
// la.g_out is (n_d, n_t, 1 + 3 nfe, ns + nq): row 0 the initial values, then the collocation points of every
// element (dynamic states first, quadrature states after them); la.ind_out receives 1.0 for a failed solve.
// PACK (small theta sets, n_t <= TILE / 2 -- the shipped scripts use 50 samples): a tile holds the whole theta set of
// TILE / n_t consecutive design points instead of one design point's slice, so the lanes are not left idle; the
// weighted indicator is summed per design point through shared memory in scenario order (deterministic).
// CHAIN (0, 2 or 3): the program's element loop is one X_PCHAIN instruction over at most CHAIN blocks (poly_chain.cuh);
// such a program holds nothing but scalar ops, INIT, LDCTRL, PCHAIN, XEND, END, and this instantiation nothing else.
// SB / SQ: compile-time shape of the chain (chain_shapes.h lists the registered ones; 0 = decoded at run time).
template <int MAXNS, int SPT, bool TRAJ = false, bool PACK = false, int CHAIN = 0, unsigned long long SB = 0ull, unsigned SQ = 0u>
#ifndef PYDSB_CHAIN_CTAS
#define PYDSB_CHAIN_CTAS 3
#endif
__global__ void __launch_bounds__(BLOCK, (SPT == 2 ? (CHAIN > 0 && SB != 0ull ? PYDSB_CHAIN_CTAS : 3) : (MAXNS <= 1 ? 5 : (MAXNS == 2 ? 4 : 2))))
    feas_kernel(const __grid_constant__ ExecImage img, const __grid_constant__ LaunchArgs la)
{
    typedef Vec<SPT> VecT;
    constexpr unsigned USP = US * SPT;                                  // bytes between consecutive slots
    constexpr int TILE = SPT * BLOCK;                                   // theta rows per tile
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);                 // [1]
    double* red = reinterpret_cast<double*>(smem + 16);                 // [BLOCK/32]
    // constant pool; CHAIN: every entry SPT-wide (one VecT load serves both scenarios) after the specials 1.0, -0.0, 0.0
    constexpr int CW = CHAIN > 0 ? SPT : 1, CPRE = CHAIN > 0 ? 3 : 1;
    double* cpool = reinterpret_cast<double*>(smem + img.off_const) + CPRE * CW;   // cpool[-1] = 0.0
    uint16_t* tabs = reinterpret_cast<uint16_t*>(smem + img.off_tabs);
    double* land = reinterpret_cast<double*>(smem + img.off_land);      // [TILE*p_dim]
    double* S = reinterpret_cast<double*>(smem + img.off_slots) + tid * SPT;   // S[slot * TILE + s]

    // ---- stage constants and reference tables in shared memory (once per CTA) ----
    for (int i = tid; i < img.n_const * CW; i += BLOCK) cpool[i] = img.consts[i / CW];
    for (int i = tid; i < img.n_tabs; i += BLOCK) tabs[i] = img.tabs[i];
    if (tid == 0) {
        for (int s = 0; s < CW; ++s) cpool[-1 * CW + s] = 0.0;
        if constexpr (CHAIN > 0)
            for (int s = 0; s < CW; ++s) { cpool[-3 * CW + s] = 1.0; cpool[-2 * CW + s] = -0.0; }
        mbar_init(&mbar[0], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const uint16_t* tab_par = tabs;
    const uint16_t* tab_x0 = tab_par + img.d_dim + img.p_dim;
    const uint16_t* tab_q0 = tab_x0 + img.ns;
    const uint16_t* tab_z = tab_q0 + img.nq;
    const uint16_t* tab_zfix = tab_z + img.nz;
    const uint16_t* tab_g = tab_zfix + img.nz;
    const uint16_t* tab_cmap = tab_g + img.ng;

    // cold-path access through 14-bit references (tape_vm.cuh): slot of scenario s, or constant pool
    auto slot = [&](int idx, int s) -> double& { return S[(size_t)idx * TILE + s]; };
    auto ref_ld = [&](unsigned ref, int s) -> double {
        if (ref & REF_CONST) return (ref == REF_NONE) ? 0.0 : cpool[(ref & REF_MASK) * CW];
        return S[(size_t)(ref & REF_MASK) * TILE + s];
    };

    char* const base = reinterpret_cast<char*>(S);
    const char* const cz = reinterpret_cast<const char*>(cpool - CPRE * CW);    // constant operands: (index + CPRE) * 8 * CW

    const double h = 1.0 / (double)img.nfe;
    const int p_dim = img.p_dim, d_dim = img.d_dim;
    // TRAJ: quadrature states at the inner collocation points need the integration matrix = inverse of the
    // differentiation matrix the Newton systems use (its last row are the quadrature weights bw)
    double aint[TRAJ ? 2 : 1][3];
    if constexpr (TRAJ) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double M[3][3], r[3] = {c == 0 ? 1.0 : 0.0, c == 1 ? 1.0 : 0.0, c == 2 ? 1.0 : 0.0};
#pragma unroll
            for (int jj = 0; jj < 3; ++jj)
#pragma unroll
                for (int kk = 0; kk < 3; ++kk) M[jj][kk] = img.adot[4 * (kk + 1) + jj + 1];
            lu_solve_inplace<3>(M, r);
            aint[0][c] = r[0];
            aint[1][c] = r[1];
        }
    }
    const int traj_w = img.ns + img.nq;                                 // TRAJ: values per time point
    const long long traj_rows = 1 + 3 * (long long)img.nfe;

    const int pack_k = PACK ? TILE / (int)la.n_t : 1;                   // PACK: design points per tile
    const long long total = PACK ? (la.n_d + pack_k - 1) / pack_k : la.n_d * (long long)la.n_tiles;
    unsigned phase = 0;

    // theta rows tile tj of a design point needs in the landing buffer (PACK: the whole set, every tile)
    auto tile_rows = [&](int tj) -> int {
        if constexpr (PACK) return (int)la.n_t;
        const long long rem = la.n_t - (long long)tj * TILE;
        return rem < TILE ? (int)rem : TILE;
    };
    auto tile_tma = [&](int rows) -> bool {
        return la.theta_tma_ok && ((rows * p_dim * (int)sizeof(double)) % 16 == 0);
    };
    auto issue = [&](int tj) {
        const int rows = tile_rows(tj);
        if (tile_tma(rows)) {
            const long long j0 = PACK ? 0 : (long long)tj * TILE;
            const unsigned bytes = (unsigned)(rows * p_dim * sizeof(double));
            mbar_expect_tx(&mbar[0], bytes);
            tma_load_1d(land, la.theta + j0 * p_dim, bytes, &mbar[0]);
        }
    };

    // tile t = (design point t / n_tiles, theta tile t % n_tiles): divided once, then advanced by the grid stride
    long long t = blockIdx.x;
    long long i_dn = PACK ? 0 : t / la.n_tiles;
    int tjn = PACK ? 0 : (int)(t % la.n_tiles);
    const int step_q = PACK ? 0 : (int)(gridDim.x / (unsigned)la.n_tiles), step_r = PACK ? 0 : (int)(gridDim.x % (unsigned)la.n_tiles);
    if (t < total && tid == 0) issue(tjn);

    unsigned long long cnt_iters = 0, cnt_flops = 0;
    unsigned cnt_warp_iters = 0, cnt_failed = 0, cnt_done = 0;

    for (; t < total; t += gridDim.x) {
        const long long i_d = PACK ? t * pack_k : i_dn;                // PACK: the tile's first design point
        const int tj = tjn;
        i_dn += step_q; tjn += step_r;                                 // the CTA's next tile
        if (tjn >= la.n_tiles) { tjn -= la.n_tiles; ++i_dn; }
        const long long j0 = PACK ? 0 : (long long)tj * TILE;
        const int rows = tile_rows(tj);
        bool valid[SPT];
        long long jj_[SPT];
        int row[SPT];                                      // row of the landing buffer
        int dl[PACK ? SPT : 1];                            // PACK: design point within the tile
        int pack_here = 1;                                 // PACK: design points in this tile
        if constexpr (PACK) {
            const long long left = la.n_d - i_d;
            pack_here = left < pack_k ? (int)left : pack_k;
            const int n_scen = pack_here * (int)la.n_t;
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                const int r = tid * SPT + s;
                valid[s] = r < n_scen;
                const int rr = valid[s] ? r : n_scen - 1;  // idle lanes replay the last scenario
                dl[s] = rr / (int)la.n_t;
                row[s] = rr - dl[s] * (int)la.n_t;
                jj_[s] = row[s];
            }
        } else {
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                const int r = tid * SPT + s;               // a thread's scenarios are consecutive theta rows
                valid[s] = r < rows;
                row[s] = valid[s] ? r : rows - 1;          // idle lanes replay the last row
                jj_[s] = j0 + row[s];
            }
        }
        // design point of the thread's scenario s
        auto id_of = [&](int s) -> long long {
            if constexpr (PACK) return i_d + dl[s];
            else return i_d;
        };

        // ---- parameters into the register file (theta from the TMA landing buffer) ----
        // slot of parameter c of scenario s (chain programs: from the descriptor, no table read)
        auto par_slot = [&](int c, int s) -> double& {
            if constexpr (CHAIN > 0) return *reinterpret_cast<double*>(reinterpret_cast<char*>(S) + c_chain.par_off[c] + s * 8);
            else return slot(tab_par[c], s);
        };
        if (tile_tma(rows)) {
            mbar_wait(&mbar[0], phase);
            phase ^= 1u;
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                const double* src = land + (size_t)row[s] * p_dim;
                int c0 = 0;
                for (; c0 + 4 <= p_dim; c0 += 4) {         // four at a time: the loads issue back to back
                    const double v0 = src[c0], v1 = src[c0 + 1], v2 = src[c0 + 2], v3 = src[c0 + 3];
                    par_slot(d_dim + c0, s) = v0; par_slot(d_dim + c0 + 1, s) = v1;
                    par_slot(d_dim + c0 + 2, s) = v2; par_slot(d_dim + c0 + 3, s) = v3;
                }
                for (int c = c0; c < p_dim; ++c) par_slot(d_dim + c, s) = src[c];
            }
        } else {
#pragma unroll
            for (int s = 0; s < SPT; ++s) {
                const double* src = la.theta + jj_[s] * p_dim;
                for (int c = 0; c < p_dim; ++c) par_slot(d_dim + c, s) = __ldg(src + c);
            }
        }
        if constexpr (PACK) {
#pragma unroll
            for (int s = 0; s < SPT; ++s)
                for (int c = 0; c < d_dim; ++c) par_slot(c, s) = __ldg(la.d + id_of(s) * d_dim + c);
        } else {
            const double* drow = la.d + i_d * d_dim;
            int c0 = 0;
            for (; c0 + 4 <= d_dim; c0 += 4) {
                const double v0 = __ldg(drow + c0), v1 = __ldg(drow + c0 + 1), v2 = __ldg(drow + c0 + 2), v3 = __ldg(drow + c0 + 3);
#pragma unroll
                for (int s = 0; s < SPT; ++s) {
                    par_slot(c0, s) = v0; par_slot(c0 + 1, s) = v1; par_slot(c0 + 2, s) = v2; par_slot(c0 + 3, s) = v3;
                }
            }
            for (int c = c0; c < d_dim; ++c) {
                const double dv = __ldg(drow + c);
#pragma unroll
                for (int s = 0; s < SPT; ++s) par_slot(c, s) = dv;
            }
        }
        __syncthreads();                                   // landing buffer and `red` are free again
        {
            const long long nt = t + gridDim.x;            // prefetch the next tile's theta rows
            if (nt < total && tid == 0) issue(tjn);
        }

        // ---- the scenarios' program ----
        bool failed[SPT], conv[SPT];
        unsigned long long my_acc[SPT];                    // Newton iterations << 40 | algorithmic flops
        double xs[SPT][MAXNS], zv[SPT][3][MAXNS];
        double s_prev[SPT];
#pragma unroll
        for (int s = 0; s < SPT; ++s) { failed[s] = false; conv[s] = false; my_acc[s] = 0; s_prev[s] = 0.0; }
        int it = 0, e = 0;
        int pcw = 0;                                       // uint4 index of the next instruction
// NS: w0 = { op, target pcw, flops, jmask | max_it << 16 }.  Stop when the step is below tolerance, or when
// the contraction rate seen so far bounds the remaining error below it:  |dz_k|^2 / |dz_{k-1}| <= rtol |X|
// (with |dz_k| <= 0.1 |dz_{k-1}|).  A non-finite step never passes the test: the scenario runs into the
// iteration cap and is reported failed.
#define X_NS_CASE(K, MASKED)                                                                                    \
    {                                                                                                           \
        const int at = pcw - 2;                                                                                 \
        pcw = at + ns_length(K);                                                                                \
        double sq[SPT], sc[SPT];                                                                                \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s)                                                         \
            if (!conv[s]) my_acc[s] += ((1ull << 40) | w0.z);                                                   \
        newton_step<K, MAXNS, MASKED, SPT>(img, at, w0.w & 0xFFFFu, base, h, xs, zv, sq, sc);                   \
        bool all = true;                                                                                        \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s) {                                                       \
            const double lim = fma(img.rtol2, sc[s], img.rtol2);                                                \
            const bool ok = (sq[s] <= lim) | ((sq[s] * sq[s] <= lim * s_prev[s]) & (sq[s] <= 0.01 * s_prev[s])); \
            s_prev[s] = sq[s];                                                                                  \
            conv[s] = conv[s] | ok;                                                                             \
            all = all & conv[s];                                                                                \
        }                                                                                                       \
        ++cnt_warp_iters;                                                                                       \
        ++it;                                                                                                   \
        if (!__all_sync(0xffffffffu, all) && it < (int)(w0.w >> 16)) pcw = (int)w0.y;                           \
        else {                                                                                                  \
            _Pragma("unroll") for (int s = 0; s < SPT; ++s) failed[s] = failed[s] | !conv[s];                   \
        }                                                                                                       \
    }                                                                                                           \
    break;
// NSL: the block is linear in its own states (its Jacobian does not depend on them): one exact step
#define X_NSL_CASE(K, MASKED)                                                                                   \
    {                                                                                                           \
        const int at = pcw - 2;                                                                                 \
        pcw = at + ns_length(K);                                                                                \
        double sq[SPT], sc[SPT];                                                                                \
        newton_step<K, MAXNS, MASKED, SPT>(img, at, w0.w & 0xFFFFu, base, h, xs, zv, sq, sc);                   \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s) {                                                       \
            my_acc[s] += ((1ull << 40) | w0.z);                                                                 \
            failed[s] = failed[s] | !(sq[s] <= DBL_MAX);                                                        \
        }                                                                                                       \
        ++cnt_warp_iters;                                                                                       \
    }                                                                                                           \
    break;
#define X_NB_CASE(K)                                                                                            \
    {                                                                                                           \
        const unsigned xo[3] = {w0.y, w0.z, w0.w};                                                              \
        newton_begin<K, MAXNS, SPT>(xo, base, xs, zv);                                                          \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s) { conv[s] = false; s_prev[s] = 0.0; }                   \
        it = 0;                                                                                                 \
    }                                                                                                           \
    break;
// POLY1: a one-state block whose right-hand side is  f = c0 + c1 x + c2 x^2  in its own state (coefficients:
// parameter / control values or functions of earlier blocks' states, fixed during this solve).  The whole
// Newton loop runs here, in registers: no dispatch, no register-file traffic until the converged points are
// stored.  Same equations, start and stop rule as NB + point ops + NS.
#ifndef PYDSB_POLY1_ZFORM
// state = the point values x_j themselves; the linear part uses all four differentiation weights (one more FMA per
// point than the increment form below, but no x_s + z_j additions: 3 fewer FP64 instructions per iteration, +2.5 % on C3)
#define POLY1_STATE_DECL double xq[SPT][3];
#define POLY1_STATE_INIT xq[s][0] = xq[s][1] = xq[s][2] = xsv.v[s];
#define POLY1_X_AND_LIN                                                                                         \
    const double x = xq[s][jj];                                                                                 \
    double lin = img.adot[jj + 1] * xsv.v[s];                                                                   \
    lin = fma(img.adot[4 + jj + 1], xq[s][0], lin);                                                             \
    lin = fma(img.adot[8 + jj + 1], xq[s][1], lin);                                                             \
    lin = fma(img.adot[12 + jj + 1], xq[s][2], lin);
#define POLY1_UPDATE                                                                                            \
    xq[s][kk] += r[kk];                                                                                         \
    const double x = xq[s][kk];
#define POLY1_X_OUT xq[s][kk]
#else
// state = the increments z_j = x_j - x_s (the rows of the differentiation matrix sum to zero)
#define POLY1_STATE_DECL double z[SPT][3];
#define POLY1_STATE_INIT z[s][0] = z[s][1] = z[s][2] = 0.0;
#define POLY1_X_AND_LIN                                                                                         \
    const double x = xsv.v[s] + z[s][jj];                                                                       \
    double lin = img.adot[4 + jj + 1] * z[s][0];                                                                \
    lin = fma(img.adot[8 + jj + 1], z[s][1], lin);                                                              \
    lin = fma(img.adot[12 + jj + 1], z[s][2], lin);
#define POLY1_UPDATE                                                                                            \
    z[s][kk] += r[kk];                                                                                          \
    const double x = xsv.v[s] + z[s][kk];
#define POLY1_X_OUT (xsv.v[s] + z[s][kk])
#endif
#define X_POLY1_CASE                                                                                            \
    {                                                                                                           \
        const uint4 w1 = c_prog[pcw - 1];                                                                       \
        const unsigned st0 = (w1.w & 1u) ? USP : 0u, st1 = (w1.w & 2u) ? USP : 0u, st2 = (w1.w & 4u) ? USP : 0u; \
        const bool linear = (w1.w & 8u) != 0;                                                                   \
        VecT c0v[3], c1v[3], c2v[3];                                                                            \
        _Pragma("unroll") for (unsigned p = 0; p < 3; ++p) {                                                    \
            c1v[p] = LDV(w1.y + p * st1);                                                                       \
            c2v[p] = LDV(w1.z + p * st2);                                                                       \
        }                                                                                                       \
        if (w1.w & 16u) {                          /* c0 = narrow * wide^2, operands in a third word */         \
            const uint4 w2 = c_prog[pcw];                                                                       \
            pcw += 1;                                                                                           \
            const VecT ka = LDV(w2.x);                                                                          \
            _Pragma("unroll") for (unsigned p = 0; p < 3; ++p) {                                                \
                const VecT cb = LDV(w2.y + p * USP);                                                            \
                _Pragma("unroll") for (int s = 0; s < SPT; ++s) c0v[p].v[s] = ka.v[s] * cb.v[s] * cb.v[s];      \
            }                                                                                                   \
        } else {                                                                                                \
            _Pragma("unroll") for (unsigned p = 0; p < 3; ++p) c0v[p] = LDV(w1.x + p * st0);                    \
        }                                                                                                       \
        const VecT xsv = LDVP(w0.y, 2);                                                                         \
        POLY1_STATE_DECL                                                                                        \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s) {                                                       \
            POLY1_STATE_INIT                                                                                    \
            conv[s] = false;                                                                                    \
            s_prev[s] = 0.0;                                                                                    \
        }                                                                                                       \
        for (it = 0;;) {                                                                                        \
            bool all = true;                                                                                    \
            _Pragma("unroll") for (int s = 0; s < SPT; ++s) {                                                   \
                if (!conv[s]) my_acc[s] += ((1ull << 40) | w0.z);                                               \
                double M[3][3], r[3];                                                                           \
                _Pragma("unroll") for (int jj = 0; jj < 3; ++jj) {                                              \
                    POLY1_X_AND_LIN                                                                             \
                    const double t = fma(c2v[jj].v[s], x, c1v[jj].v[s]);                                        \
                    const double fx = fma(t, x, c0v[jj].v[s]);                                                  \
                    const double jx = fma(c2v[jj].v[s], x, t);                                                  \
                    r[jj] = fma(h, fx, -lin);                                                                   \
                    _Pragma("unroll") for (int kk = 0; kk < 3; ++kk) M[jj][kk] = img.adot[4 * (kk + 1) + jj + 1]; \
                    M[jj][jj] = fma(-h, jx, M[jj][jj]);                                                         \
                }                                                                                               \
                lu_solve_inplace<3>(M, r);                                                                      \
                double sq = 0.0, sc = 0.0;                                                                      \
                _Pragma("unroll") for (int kk = 0; kk < 3; ++kk) {                                              \
                    POLY1_UPDATE                                                                                \
                    sq = fma(r[kk], r[kk], sq);                                                                 \
                    sc = fma(x, x, sc);                                                                         \
                }                                                                                               \
                const double lim = fma(img.rtol2, sc, img.rtol2);                                               \
                const bool ok = linear ? (sq <= DBL_MAX)                                                        \
                                       : ((sq <= lim) | ((sq * sq <= lim * s_prev[s]) & (sq <= 0.01 * s_prev[s]))); \
                s_prev[s] = sq;                                                                                 \
                conv[s] = conv[s] | ok;                                                                         \
                all = all & conv[s];                                                                            \
            }                                                                                                   \
            ++cnt_warp_iters;                                                                                   \
            ++it;                                                                                               \
            if (linear || __all_sync(0xffffffffu, all) || it >= (int)w0.w) break;                               \
        }                                                                                                       \
        VecT xo[3];                                                                                             \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s) {                                                       \
            failed[s] = failed[s] | !conv[s];                                                                   \
            _Pragma("unroll") for (int kk = 0; kk < 3; ++kk) xo[kk].v[s] = POLY1_X_OUT;                         \
        }                                                                                                       \
        STVP(w0.y, 0, xo[0]); STVP(w0.y, 1, xo[1]); STVP(w0.y, 2, xo[2]);                                       \
    }                                                                                                           \
    break;
// INIT: initial states (seeded at collocation point 2), quadratures
#define X_INIT_BODY                                                                                             \
    {                                                                                                           \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s) {                                                       \
            double x0v[MAX_STATES];                                                                             \
            for (int n = 0; n < img.ns; ++n) x0v[n] = ref_ld(tab_x0[n], s);                                     \
            for (int n = 0; n < img.ns; ++n) slot(img.x_slot + 3 * n + 2, s) = x0v[n];                          \
            for (int q = 0; q < img.nq; ++q) slot(img.qe_slot + q, s) = ref_ld(tab_q0[q], s);                   \
            if (img.zero_slot) slot(img.zero_slot, s) = 0.0;                                                    \
            if constexpr (TRAJ) {                                                                               \
                if (valid[s]) {                                                                                 \
                    double* o = la.g_out + (id_of(s) * la.n_t + jj_[s]) * traj_rows * traj_w;                   \
                    for (int n = 0; n < img.ns; ++n) o[n] = x0v[n];                                             \
                    for (int qn = 0; qn < img.nq; ++qn) o[img.ns + qn] = slot(img.qe_slot + qn, s);             \
                }                                                                                               \
            }                                                                                                   \
        }                                                                                                       \
    }
// LDCTRL: control w0.z of finite element e -> slot w0.y
#define X_LDCTRL_BODY                                                                                           \
    {                                                                                                           \
        VecT zv_;                                                                                               \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s) zv_.v[s] = ctrl_ld(e, (int)w0.z, s);                    \
        STV(w0.y, zv_);                                                                                         \
    }
// XEND: end-point states for the CQA tape
#define X_XEND_BODY                                                                                             \
    {                                                                                                           \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s)                                                         \
            for (int n = 0; n < img.ns; ++n) slot(img.xe_slot + n, s) = slot(img.x_slot + 3 * n + 2, s);        \
    }
        // value of control c on finite element el for the thread's scenario s (z_mode: where the controls come from)
        auto ctrl_ld = [&](int el, int c, int s) -> double {
            const int zi = tab_cmap[el * img.nce + c];
            if (la.z_mode == 0 || tab_zfix[zi]) {
                // chain programs: the assembler renumbers the units, the reference table of the lowering does not apply
                if constexpr (CHAIN > 0)
                    return *reinterpret_cast<const double*>((((c_chain.z_const >> zi) & 1u) ? cz : base) + c_chain.z_off[zi] + s * 8);
                else return ref_ld(tab_z[zi], s);
            }
            if (la.z_mode == 1) return __ldg(la.z + zi);
            if (la.z_mode == 2) return __ldg(la.z + id_of(s) * img.nz + zi);
            return __ldg(la.z + (id_of(s) * la.n_t + jj_[s]) * img.nz + zi);
        };
        if constexpr (CHAIN > 0) {
            // A chain program has a fixed shape: two counted loops over scalar ops around the fused element loop.  Operands
            // of the scalar ops: a slot of this thread or an entry of the constant pool (staged SPT-wide, so both are one
            // VecT load from a selected base); the FMA-class ops (MOV NEG ADD SUB MUL FMA FMS FNMA SQR) are ONE op, FMAG:
            // r = fma(+-a, b, +-c) with 1.0 / -0.0 from the pool -- bit-identical to the individual ops, no dispatch.
            // a run of n FMAG ops: nothing but fetch, three operand loads, two sign flips, the FMA and the store
            auto fmag_run = [&](int pc, const int n) -> int {
                for (int i = 0; i < n; ++i, pc += 2) {
                    const uint4 w0 = c_prog[pc], w1 = c_prog[pc + 1];
                    const unsigned fl = w1.y;
                    const VecT av = *reinterpret_cast<const VecT*>(((fl & 1u) ? cz : base) + w0.z);
                    const VecT bv = *reinterpret_cast<const VecT*>(((fl & 2u) ? cz : base) + w0.w);
                    const VecT cv = *reinterpret_cast<const VecT*>(((fl & 4u) ? cz : base) + w1.x);
                    const int sa = (int)((fl & 8u) << 28), sc = (int)((fl & 16u) << 27);        // sign bits to flip
                    VecT rv;
#pragma unroll
                    for (int s = 0; s < SPT; ++s) {
                        const double a = __hiloint2double(__double2hiint(av.v[s]) ^ sa, __double2loint(av.v[s]));
                        const double c = __hiloint2double(__double2hiint(cv.v[s]) ^ sc, __double2loint(cv.v[s]));
                        rv.v[s] = fma(a, bv.v[s], c);
                    }
                    STV(w0.y, rv);
                }
                return pc;
            };
            // [pc, pc_end): run0 FMAG ops, then { one other op, the FMAG run that follows it (length in its w1.w) } ...
            auto scalar_run = [&](int pc, const int pc_end, const int run0) {
                pc = fmag_run(pc, run0);
                while (pc < pc_end) {
                    const uint4 w0 = c_prog[pc], w1 = c_prog[pc + 1];
                    const unsigned fl = w1.y;
                    const unsigned op = w0.x;
                    VecT rv;
                    if (op == X_LDCTRL) {
#pragma unroll
                        for (int s = 0; s < SPT; ++s) rv.v[s] = ctrl_ld(0, (int)w1.z, s);      // element-invariant controls only
                    } else {
                        const VecT av = *reinterpret_cast<const VecT*>(((fl & 1u) ? cz : base) + w0.z);
                        const VecT bv = *reinterpret_cast<const VecT*>(((fl & 2u) ? cz : base) + w0.w);
                        if (op == X_COLD1 && w1.z == OP_EXP) {
#pragma unroll
                            for (int s = 0; s < SPT; ++s) rv.v[s] = exp(av.v[s]);
                        } else {
#pragma unroll
                            for (int s = 0; s < SPT; ++s) {
                                const double a = av.v[s], b = bv.v[s];
                                double r;
                                switch (op) {
                                // MUFU seed + Newton steps and one residual correction: within 1 ulp of a / b, no slow path
                                case X_DIV1: { const double rb = fast_rcp(b), q = a * rb; r = fma(fma(-b, q, a), rb, q); break; }
                                case X_RCP1: { const double ra = fast_rcp(a); r = fma(fma(-a, ra, 1.0), ra, ra); break; }
                                case X_ABS1: r = fabs(a); break;
                                case X_MIN1: r = fmin(a, b); break;
                                case X_MAX1: r = fmax(a, b); break;
                                default: r = cold_op(w1.z, a, b); break;       // X_COLD1
                                }
                                rv.v[s] = r;
                            }
                        }
                    }
                    STV(w0.y, rv);
                    pc = fmag_run(pc + 2, (int)w1.w);
                }
            };
            scalar_run(0, (int)c_chain.pre_end, (int)c_chain.pre_run0);
            poly_chain<SPT, CHAIN, SB, SQ>(img, base, cz, h, ctrl_ld, failed, my_acc, cnt_warp_iters);
            scalar_run((int)c_chain.g_begin, (int)c_chain.g_end, (int)c_chain.g_run0);
        } else
        for (;;) {
            const uint4 w0 = c_prog[pcw];
            pcw += 2;
            const unsigned op = w0.x;
            if (op < X_TIER1) {
                switch (op) {
                case X_POLY1: X_POLY1_CASE
                case X_MUL_NW: X_BIN(N, W, a * b)
                case X_NS1: X_NS_CASE(1, false)
                case X_ELEM_END: {
                    // quadratures  x_q += h * sum_j bw_j fq(x_j)  at the converged points, then the next element
                    const int at = pcw - 2;
                    const int nquad = (int)w0.z;
                    if constexpr (TRAJ) {
                        // converged collocation points of the dynamic states; quadrature states without an entry
                        // below (zero integrand) keep their value
#pragma unroll
                        for (int s = 0; s < SPT; ++s) {
                            if (!valid[s]) continue;
                            double* o = la.g_out + ((id_of(s) * la.n_t + jj_[s]) * traj_rows + 1 + 3 * e) * traj_w;
                            for (int p = 0; p < 3; ++p) {
                                for (int n = 0; n < img.ns; ++n) o[p * traj_w + n] = slot(img.x_slot + 3 * n + p, s);
                                for (int qn = 0; qn < img.nq; ++qn) o[p * traj_w + img.ns + qn] = slot(img.qe_slot + qn, s);
                            }
                        }
                    }
                    for (int qi = 0; qi < nquad; ++qi) {
                        const uint4 qw = c_prog[at + 1 + 2 * qi];
                        const uint4 qs = c_prog[at + 2 + 2 * qi];      // { scale offset, has scale }
                        const VecT f0 = LDV(qw.y), f1 = LDV(qw.z), f2 = LDV(qw.w);
                        VecT q = LDV(qw.x), acc;
#pragma unroll
                        for (int s = 0; s < SPT; ++s) {
                            acc.v[s] = img.bw[0] * f0.v[s];
                            acc.v[s] = fma(img.bw[1], f1.v[s], acc.v[s]);
                            acc.v[s] = fma(img.bw[2], f2.v[s], acc.v[s]);
                        }
                        VecT sc;
#pragma unroll
                        for (int s = 0; s < SPT; ++s) sc.v[s] = 1.0;
                        if (qs.y) {                                    // uniform branch: integrand = scale * (wide value)
                            sc = LDV(qs.x);
#pragma unroll
                            for (int s = 0; s < SPT; ++s) acc.v[s] *= sc.v[s];
                        }
                        if constexpr (TRAJ) {
                            const int qn = (int)(qw.x / USP) - img.qe_slot;
#pragma unroll
                            for (int s = 0; s < SPT; ++s) {
                                if (!valid[s]) continue;
                                double* o = la.g_out + ((id_of(s) * la.n_t + jj_[s]) * traj_rows + 1 + 3 * e) * traj_w + img.ns + qn;
                                for (int p = 0; p < 2; ++p) {
                                    double a = aint[p][0] * f0.v[s];
                                    a = fma(aint[p][1], f1.v[s], a);
                                    a = fma(aint[p][2], f2.v[s], a);
                                    o[p * traj_w] = fma(h, a * sc.v[s], q.v[s]);
                                }
                                o[2 * traj_w] = fma(h, acc.v[s], q.v[s]);
                            }
                        }
#pragma unroll
                        for (int s = 0; s < SPT; ++s) q.v[s] = fma(h, acc.v[s], q.v[s]);
                        STV(qw.x, q);
                    }
                    pcw = (++e < img.nfe) ? (int)w0.y : at + 1 + 2 * nquad;
                    break;
                }
                }
            } else if (op < X_TIER2) {
                switch (op) {
                case X_FMA_NWN: X_TER(N, W, N, fma(a, b, c))
                case X_FMA_NWW: X_TER(N, W, W, fma(a, b, c))
                case X_SQR_W: X_UN(W, a * a)
                case X_FMS_NWW: X_TER(N, W, W, fma(a, b, -c))
                case X_MUL_WW: X_BIN(W, W, a * b)
                case X_MULSQ_NW: X_BIN(N, W, a * b * b)
                case X_MUL1: X1_BIN(a * b)
                case X_FMA1: X1_TER(fma(a, b, c))
                case X_ADD1: X1_BIN(a + b)
                case X_NS1L: X_NSL_CASE(1, false)
                case X_NS2:
                    if constexpr (MAXNS >= 2) X_NS_CASE(2, true)
                    break;
                case X_NB1: X_NB_CASE(1)
                }
            } else {
                switch (op) {
                case X_ADD_NW: X_BIN(N, W, a + b)
                case X_FMA_WWN: X_TER(W, W, N, fma(a, b, c))
                case X_FMA_WWW: X_TER(W, W, W, fma(a, b, c))
                case X_FMS_NWN: X_TER(N, W, N, fma(a, b, -c))
                case X_FMS_WWN: X_TER(W, W, N, fma(a, b, -c))
                case X_FMS_WWW: X_TER(W, W, W, fma(a, b, -c))
                case X_FNMA_NWN: X_TER(N, W, N, fma(-a, b, c))
                case X_FNMA_NWW: X_TER(N, W, W, fma(-a, b, c))
                case X_FNMA_WWN: X_TER(W, W, N, fma(-a, b, c))
                case X_FNMA_WWW: X_TER(W, W, W, fma(-a, b, c))
                case X_ADD_WW: X_BIN(W, W, a + b)
                case X_SUB_WW: X_BIN(W, W, a - b)
                case X_SUB_NW: X_BIN(N, W, a - b)
                case X_SUB_WN: X_BIN(W, N, a - b)
                case X_DIV_NW: X_BIN(N, W, a / b)
                case X_DIV_WN: X_BIN(W, N, a / b)
                case X_DIV_WW: X_BIN(W, W, a / b)
                case X_NEG_W: X_UN(W, -a)
                case X_RCP_W: X_UN(W, 1.0 / a)
                case X_ABS_W: X_UN(W, fabs(a))
                case X_MOV_W: X_UN(W, a)
                case X_BC_N: X_UN(N, a)
                case X_MIN_NW: X_BIN(N, W, fmin(a, b))
                case X_MIN_WW: X_BIN(W, W, fmin(a, b))
                case X_MAX_NW: X_BIN(N, W, fmax(a, b))
                case X_MAX_WW: X_BIN(W, W, fmax(a, b))
                case X_COLD_W: {                           // transcendental, 3-wide; flag bit 0/1: a/b is wide
                    const uint4 w1 = c_prog[pcw - 1];
                    const unsigned sa = (w1.y & 1u) ? USP : 0u, sb = (w1.y & 2u) ? USP : 0u;
#pragma unroll 1
                    for (unsigned p = 0; p < 3; ++p) {
                        const VecT a = LDV(w0.z + p * sa), b = LDV(w0.w + p * sb);
                        VecT r;
#pragma unroll
                        for (int s = 0; s < SPT; ++s) r.v[s] = cold_op(w1.z, a.v[s], b.v[s]);
                        STV(w0.y + p * USP, r);
                    }
                    break;
                }
                case X_FMS1: X1_TER(fma(a, b, -c))
                case X_FNMA1: X1_TER(fma(-a, b, c))
                case X_SUB1: X1_BIN(a - b)
                case X_DIV1: X1_BIN(a / b)
                case X_SQR1: X1_UN(a * a)
                case X_NEG1: X1_UN(-a)
                case X_RCP1: X1_UN(1.0 / a)
                case X_ABS1: X1_UN(fabs(a))
                case X_MOV1: X1_UN(a)
                case X_MIN1: X1_BIN(fmin(a, b))
                case X_MAX1: X1_BIN(fmax(a, b))
                case X_COLD1: X1_BIN(cold_op(w1.z, a, b))
                case X_INIT: X_INIT_BODY break;
                case X_LDCTRL: X_LDCTRL_BODY break;
                case X_NB2:
                    if constexpr (MAXNS >= 2) X_NB_CASE(2)
                    break;
                case X_NB3:
                    if constexpr (MAXNS >= 3) X_NB_CASE(3)
                    break;
                case X_NS2L:
                    if constexpr (MAXNS >= 2) X_NSL_CASE(2, true)
                    break;
                case X_NS3L:
                    if constexpr (MAXNS >= 3) X_NSL_CASE(3, true)
                    break;
                case X_NS3:
                    if constexpr (MAXNS >= 3) X_NS_CASE(3, true)
                    break;
                case X_XEND: X_XEND_BODY break;
                default: break;
                }
                if (op == X_END) break;
            }
        }
#undef X_POLY1_CASE
#undef X_INIT_BODY
#undef X_LDCTRL_BODY
#undef X_XEND_BODY
#undef X_NS_CASE
#undef X_NSL_CASE
#undef X_NB_CASE

        // ---- indicator, outputs ----
        double wv = 0.0;
#pragma unroll
        for (int s = 0; s < SPT; ++s) {
            bool feasible = !failed[s];
            const long long j = jj_[s];
            for (int c = 0; c < img.ng; ++c) {
                double gv;
                if constexpr (CHAIN > 0) gv = *reinterpret_cast<const double*>(base + c_chain.g_off[c] + s * 8);
                else gv = ref_ld(tab_g[c], s);
                if (failed[s]) gv = __longlong_as_double(0x7ff8000000000000LL);
                feasible = feasible && (gv <= la.gtol[c]);
                if constexpr (!TRAJ)
                    if (la.g_out && valid[s]) la.g_out[(id_of(s) * la.n_t + j) * img.ng + c] = gv;
            }
            if (la.ind_out && valid[s]) la.ind_out[id_of(s) * la.n_t + j] = TRAJ ? (failed[s] ? 1.0 : 0.0) : (feasible ? -0.0 : -1.0);
            const double w1 = (valid[s] && feasible) ? (la.w ? __ldg(la.w + j) : 1.0 / (double)la.n_t) : 0.0;
            if constexpr (PACK) slot(0, s) = w1;           // the program is over: the first register-file unit is scratch
            else wv += w1;
            if (valid[s]) {
                cnt_iters += my_acc[s] >> 40;
                cnt_flops += my_acc[s] & ((1ull << 40) - 1);
                cnt_failed += failed[s] ? 1u : 0u;
                cnt_done += 1u;
            }
        }
        if constexpr (PACK) {
            // per design point: its n_t weighted indicators in scenario order, one thread each
            __syncthreads();
            for (int q = tid; q < pack_here; q += BLOCK) {
                const double* wr = reinterpret_cast<const double*>(smem + img.off_slots) + q * (int)la.n_t;
                double sum = 0.0;
                for (int j = 0; j < (int)la.n_t; ++j) sum += wr[j];
                la.partial[i_d + q] = sum;
            }
            __syncthreads();                               // before the next tile's parameters overwrite the unit
        } else {
            // one partial per warp (reduce_partials_kernel adds them in fixed order): no CTA barrier, no serial tail
            wv = warp_sum(wv);
            if (lane == 0) la.partial[t * (BLOCK / 32) + warp] = wv;
        }
    }

    // ---- counters: one atomic per warp ----
    if (la.counters) {
        const unsigned long long it_sum = warp_sum_u64(cnt_iters);
        const unsigned long long fl_sum = warp_sum_u64(cnt_flops);
        const unsigned wi = warp_sum_u32(cnt_warp_iters);
        const unsigned fl = warp_sum_u32(cnt_failed);
        const unsigned dn = warp_sum_u32(cnt_done);
        if (lane == 0) {
            atomicAdd(&la.counters[0], (unsigned long long)dn);
            atomicAdd(&la.counters[1], it_sum);
            atomicAdd(&la.counters[2], (unsigned long long)wi * SPT);
            atomicAdd(&la.counters[3], (unsigned long long)fl);
            atomicAdd(&la.counters[5], fl_sum);
        }
    }
}

}  // namespace pydsb
