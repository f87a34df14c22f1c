// Batched control (recourse) optimisation: the GPU replacement for the NLP solve the reference
// hands to GAMS/CONOPT once per design point (reference pyds/solver.py:26-55; the model is the
// big-M relaxation of pyds/utils.py:7-51).  Eliminating the relaxed indicators
// leaves, per group of scenarios that share a control vector z,
//
//     min_{z in box}  Phi(z) = sum_s w_s max(0, max_k g_ks(z) / M_k)
//
// minimised by projected Levenberg-Marquardt steps on the log-sum-exp smoothing
// phi_mu(c) = mu log(1 + sum_k exp(c_k/mu)), mu -> mu_min (the primal form of the barrier an
// interior-point method would use), with a Gauss-Newton Hessian built from forward control
// sensitivities.  One iteration = two launches:
//
//   sens_kernel<NS>   thread per scenario: state solve as in feas_kernel, plus per finite element the
//                     sensitivity systems  M dX_c = -dR/dx_prev S_c - dR/dz_c  solved with the LU
//                     factors of the converged Newton matrix (register resident); then c_k, grad c_k,
//                     softmax weights, and the per-group sums {Phi_mu, Phi, grad, H} reduced with warp
//                     shuffles (shared mode) or written per scenario (per-scenario mode)
//   lm_update_kernel  thread per group: accept / reject, damping and mu schedule, active-set
//                     projected step by Cholesky (nz <= 16), convergence mask.
//
// Groups: shared mode -> one group per design point (the reference's two-stage formulation: stage-0
// F_in shared by every theta); per-scenario mode -> one group per (d, theta).
#pragma once
#include "xvm_ops.cuh"

namespace pydsb {

#define PT_LD(off) (*reinterpret_cast<const double*>(base + (off)))

// The data ops [pcw, pcw_end) of the sensitivity program: same pre-decoded format and op bodies as feas_kernel
// (uniform fetch, pattern-specialised point ops).  Inlined at its four call sites: as an out-of-line function the
// program counter arrives in a vector register and the fetch, dispatch and operand addresses leave the uniform datapath.
__device__ __forceinline__ void run_ops(char* __restrict__ base, const char* __restrict__ cz, int pcw, int pcw_end)
{
    typedef Vec<1> VecT;
    constexpr int SPT = 1;
    constexpr unsigned USP = US;
    const uint4* const c_prog = c_prog_s;
    while (pcw < pcw_end) {
        const uint4 w0 = c_prog[pcw];
        pcw += 2;
        if (w0.x < X_TIER2) {
            switch (w0.x) {
            X_DATA_OP_CASES_HOT
            default: break;
            }
        } else {
            switch (w0.x) {
            X_DATA_OP_CASES_REST
            default: break;
            }
        }
    }
}

constexpr int MAX_NZ = 16;

struct RecourseArgs {
    const double* d;
    const double* theta;
    const double* w;           // may be null -> 1/n_t (shared) ; per-scenario mode uses weight 1
    long long n_d, n_t;
    long long n_d_total;       // design points of the whole batch (> n_d when the global group spans several processes)
    int n_tiles;
    int per_scenario;          // 0: group = design point, 1: group = scenario, 2: one group for the whole grid
    // per-group optimiser state
    double* z_try;             // (G, nz)   evaluated by sens_kernel
    const int* status;         // (G)       0 = running
    const double* mu;          // (G)
    double* acc;               // shared: (n_d, n_tiles, NV) tile partials ; per-scenario: (G, NV)
    unsigned long long* counters;
    int sens_slot;             // first unit of the per-thread sensitivity block (after the program's units)
    // per-scenario mode: the scenarios still running, compacted by lm_update_kernel after every iteration so
    // that finished scenarios cost nothing (null list = every scenario, the first iteration)
    const long long* active;
    const int* n_active;
    double mu_min;             // floor of the smoothing continuation
    int n_stages;              // shared modes: stages of the continuation whose statistics one evaluation emits
};

__host__ __device__ inline int recourse_nv(int nz) { return 2 + nz + nz * (nz + 1) / 2; }
// Smoothing continuation mu_0 > mu_1 > ... (mu_{s+1} = max(mu_s / 10, mu_min)): one state + sensitivity solve serves every
// stage that is evaluated at the same controls (the c_k and their gradients do not depend on mu).  Shared modes: the
// kernel emits the reduced statistics of up to RECOURSE_STAGES stages per evaluation; per-scenario mode: it stores the
// scenario's own {failed, c_k, grad c_k} (recourse_nvp values) and lm_update_kernel evaluates any stage itself.
constexpr int RECOURSE_STAGES = 6;
__host__ __device__ inline int recourse_nvp(int ng, int nz) { return 1 + ng + ng * nz; }

// smoothed hinge of one scenario: phi = m + mu log(e^{-m/mu} + sum_k e^{(c_k-m)/mu}) with m = max(0, max_k c_k), softmax
// weights p_k; a failed state solve counts as indicator 1 with no slope
__device__ __forceinline__ void hinge_point(int ng, const double* ck, double cmax, bool failed, double mu, double* pk,
                                            double& phi, double& raw)
{
    if (failed) {
        phi = raw = 1.0;
        for (int k = 0; k < ng; ++k) pk[k] = 0.0;
        return;
    }
    double Z = exp(-cmax / mu);
    for (int k = 0; k < ng; ++k) { pk[k] = exp((ck[k] - cmax) / mu); Z += pk[k]; }
    const double zi = 1.0 / Z;
    for (int k = 0; k < ng; ++k) pk[k] *= zi;
    phi = cmax + mu * log(Z);
    raw = cmax;
}
// rows (of nz units each) of the per-thread sensitivity block: Sx (ns) and Sq (nq) while the elements are
// integrated, then Dc (ng) + one spare row
__host__ __device__ inline int recourse_sens_rows(int ns, int nq, int ng) { return ns + nq > ng + 1 ? ns + nq : ng + 1; }

// What one evaluation hands to the optimiser, shared by both evaluators (sens_kernel here, rsens_kernel in
// recourse_chain.cuh): per-scenario mode the scenario's own {failed, c_k, grad c_k}; shared modes the reduced statistics
// {Phi_mu, Phi, grad, Gauss-Newton H} of RECOURSE_STAGES smoothing stages, summed over the tile in a fixed order.
// Dc[(k * nz + c) * BLOCK] = d c_k / d z_c of this thread's scenario (shared memory, one spare row of nz after the ng
// rows); wsum: scratch_rows (<= 32) rows of BLOCK doubles for the sums over the tile, may overlay the program's slots.
__device__ __forceinline__ void recourse_emit(const RecourseArgs& ra, const int nz, const int ng, const int NV, const bool active,
                                              const bool failed, const double (&ck)[8], const double cmax, double* const Dc,
                                              double* const wsum, const int scratch_rows, const long long grp, const long long j,
                                              const long long t, const unsigned my_iters, const bool count = true)
{
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const bool per_scen = ra.per_scenario == 1;
    if (per_scen) {
        // the scenario's own c_k and grad c_k: lm_update_kernel evaluates the smoothed hinge for any mu
        if (active) {
            double* out = ra.acc + grp * recourse_nvp(ng, nz);
            out[0] = failed ? 1.0 : 0.0;
            for (int k = 0; k < ng; ++k) out[1 + k] = ck[k];
            for (int k = 0; k < ng; ++k)
                for (int c = 0; c < nz; ++c) out[1 + ng + k * nz + c] = Dc[(size_t)(k * nz + c) * BLOCK];
        }
    } else {
        double wv = 0.0;
        if (active) {
            wv = ra.w ? __ldg(ra.w + j) : 1.0 / (double)ra.n_t;
            if (ra.per_scenario == 2) wv /= (double)ra.n_d_total;   // the whole grid is one sample average
        }
        // v_c = sum_k p_k dc_kc kept in the spare row after Dc
        double* const Vv = Dc + (size_t)ng * nz * BLOCK;
        double mu = __ldcg(ra.mu + grp);                         // (rfused_kernel: updated by the group's thread 0 in between)
        const int SNV = ra.n_stages * NV;
        // Sum of every value over the tile's scenarios: the threads write their values side by side into rows of the
        // scratch (row v = value v of all threads, skewed by 4 v columns so that both the writes and the reads below are
        // conflict-free); every `scratch_rows` values the rows are summed -- four threads per row, each a quarter of the
        // columns in a fixed order, combined by two shuffles -- and stored.  (A warp_sum per value costs five dependent
        // shuffle + add rounds for each of the stages * NV values: that was most of the time of one evaluation.)
        const int cnt = __syncthreads_count(active);             // (also: every thread is done with the program's slots)
        // shared modes: the scenarios of a tile sit in the threads [0, cnt); anything else sums all columns
        const int n_cols = __syncthreads_and(active == (tid < cnt)) ? cnt : BLOCK;
        double* const buf = wsum;
        int row = 0, out0 = 0;
        auto flush = [&]() {
            __syncthreads();
            const int r = warp * 8 + (lane >> 2), part = lane & 3;
            double s0 = 0.0, s1 = 0.0;
            if (r < row) {
                const double* const rowp = buf + (size_t)r * BLOCK;
                int c = part;
                for (; c + 4 < n_cols; c += 8) {
                    s0 += rowp[(c + 4 * r) & (BLOCK - 1)];
                    s1 += rowp[(c + 4 + 4 * r) & (BLOCK - 1)];
                }
                if (c < n_cols) s0 += rowp[(c + 4 * r) & (BLOCK - 1)];
            }
            double sm = s0 + s1;
            sm += __shfl_xor_sync(0xffffffffu, sm, 1);
            sm += __shfl_xor_sync(0xffffffffu, sm, 2);
            if (r < row && part == 0) ra.acc[t * SNV + out0 + r] = sm;   // (i_d, tile) == blockIdx.x in the shared modes
            out0 += row;
            row = 0;
            __syncthreads();
        };
        auto emit = [&](double val) {                            // values in the order [stage][Phi_mu, Phi, grad, H packed]
            buf[(size_t)row * BLOCK + ((tid + 4 * row) & (BLOCK - 1))] = val;
            if (++row == scratch_rows) flush();
        };
        constexpr int NZS = 8, NVS = 2 + NZS + NZS * (NZS + 1) / 2, HALF = NVS / 2;   // the static path: nz = 8, 46 values
        const bool static_path = nz == NZS && scratch_rows >= HALF;
        for (int st = 0; st < ra.n_stages; ++st) {
            double pk[8];
            double phi, raw;
            hinge_point(ng, ck, cmax, failed, mu, pk, phi, raw);
            if (static_path) {
                // nz = 8 (the shipped problems): everything of a stage in registers with static indices -- the gradient
                // rows are read once per constraint instead of once per Hessian entry -- and two flushes of 23 rows
                double V[NZS], Hh[NZS * (NZS + 1) / 2];
#pragma unroll
                for (int c = 0; c < NZS; ++c) V[c] = 0.0;
#pragma unroll
                for (int i = 0; i < NZS * (NZS + 1) / 2; ++i) Hh[i] = 0.0;
#pragma unroll 1
                for (int k = 0; k < ng; ++k) {
                    const double* const Dk = Dc + (size_t)k * NZS * BLOCK;
                    double D[NZS];
#pragma unroll
                    for (int c = 0; c < NZS; ++c) D[c] = Dk[(size_t)c * BLOCK];
                    const double pkk = pk[k];
                    int idx = 0;
#pragma unroll
                    for (int a = 0; a < NZS; ++a) {
                        V[a] = fma(pkk, D[a], V[a]);
                        const double pa = pkk * D[a];
#pragma unroll
                        for (int b = a; b < NZS; ++b, ++idx) Hh[idx] = fma(pa, D[b], Hh[idx]);
                    }
                }
                const double wmu = wv / mu;
                {
                    int idx = 0;
#pragma unroll
                    for (int a = 0; a < NZS; ++a)
#pragma unroll
                        for (int b = a; b < NZS; ++b, ++idx) Hh[idx] = wmu * fma(-V[a], V[b], Hh[idx]);
                }
                if (row) flush();
                // value v of the stage: Phi_mu, Phi, grad (8), H packed (36); rows [0, HALF) of the scratch, twice
                auto value = [&](int v) -> double {                 // (v is a compile-time constant after unrolling)
                    if (v == 0) return wv * phi;
                    if (v == 1) return wv * raw;
                    if (v < 2 + NZS) return wv * V[v - 2];
                    return Hh[v - 2 - NZS];
                };
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int i = 0; i < HALF; ++i) buf[(size_t)i * BLOCK + ((tid + 4 * i) & (BLOCK - 1))] = value(half * HALF + i);
                    row = HALF;
                    flush();
                }
                mu = fmax(mu / 10.0, ra.mu_min);
                continue;
            }
            for (int c = 0; c < nz; ++c) {
                double v = 0.0;
                for (int k = 0; k < ng; ++k) v = fma(pk[k], Dc[(size_t)(k * nz + c) * BLOCK], v);
                Vv[(size_t)c * BLOCK] = v;
            }
            const double wmu = wv / mu;
            emit(wv * phi);
            emit(wv * raw);
            for (int c = 0; c < nz; ++c) emit(wv * Vv[(size_t)c * BLOCK]);
            for (int a = 0; a < nz; ++a)
                for (int b = a; b < nz; ++b) {
                    double hv = 0.0;
                    for (int k = 0; k < ng; ++k)
                        hv = fma(pk[k] * Dc[(size_t)(k * nz + a) * BLOCK], Dc[(size_t)(k * nz + b) * BLOCK], hv);
                    hv = fma(-Vv[(size_t)a * BLOCK], Vv[(size_t)b * BLOCK], hv);
                    emit(wmu * hv);
                }
            mu = fmax(mu / 10.0, ra.mu_min);
        }
        if (row) flush();
    }
    if (ra.counters && count) {
        const unsigned it_sum = warp_sum_u32(active ? my_iters : 0u);
        const unsigned dn = warp_sum_u32(active ? 1u : 0u);
        const unsigned fl = warp_sum_u32((active && failed) ? 1u : 0u);
        if (lane == 0) {
            atomicAdd(&ra.counters[0], (unsigned long long)dn);
            atomicAdd(&ra.counters[1], (unsigned long long)it_sum);
            atomicAdd(&ra.counters[3], (unsigned long long)fl);
        }
    }
}

// TRI: the Jacobian df/dx is lower triangular in the state order (every state depends on itself and on earlier
// states only -- A -> B -> C kinetics): the 3*NS Newton system is solved by forward substitution over NS 3x3 systems
// (the same Newton step as the dense LU, a third of the flops for NS = 2) and so are the sensitivity systems.
template <int NS, bool TRI>
__global__ void __launch_bounds__(BLOCK, (NS <= 2 ? 3 : 1))
    sens_kernel(const __grid_constant__ ExecImage img, const __grid_constant__ RecourseArgs ra)
{
    constexpr int N = 3 * NS;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    double* cpool = reinterpret_cast<double*>(smem + img.off_const) + 1;
    uint16_t* tabs = reinterpret_cast<uint16_t*>(smem + img.off_tabs);
    // [BLOCK/32][n_stages][NV] cross-warp scratch of the shared modes: overlays the program's own slots, which are
    // dead once the CQA gradients are in the sensitivity block (needs (BLOCK/32) * n_stages * NV <= n_units * BLOCK)
    double* wsum = reinterpret_cast<double*>(smem + img.off_slots);
    double* S = reinterpret_cast<double*>(smem + img.off_slots) + tid;

    for (int i = tid; i < img.n_const; i += BLOCK) cpool[i] = img.consts[i];
    for (int i = tid; i < img.n_tabs; i += BLOCK) tabs[i] = img.tabs[i];
    if (tid == 0) cpool[-1] = 0.0;
    __syncthreads();

    const uint16_t* tab_par = tabs;
    const uint16_t* tab_x0 = tab_par + img.d_dim + img.p_dim;
    const uint16_t* tab_q0 = tab_x0 + img.ns;
    const uint16_t* tab_z = tab_q0 + img.nq;
    const uint16_t* tab_zfix = tab_z + img.nz;
    const uint16_t* tab_g = tab_zfix + img.nz;
    const uint16_t* tab_cmap = tab_g + img.ng;
    const uint16_t* tab_gx = tab_cmap + img.nfe * img.nce;
    const uint16_t* tab_gq = tab_gx + img.ng * img.ns;
    const uint16_t* tab_live = tab_gq + img.ng * img.nq;       // [nfe] bit c: control c acted in an element <= e

    VM vm;
    vm.S = S;
    vm.cpool = cpool;
    char* const base = reinterpret_cast<char*>(S);
    const char* const cz = reinterpret_cast<const char*>(cpool - 1);    // constant operands: (index + 1) * 8
    double* const Xs = S + img.x_slot * BLOCK;
    const int nz = img.nz, nq = img.nq, ng = img.ng;
    double* const Sx = S + (size_t)ra.sens_slot * BLOCK;       // Sx[n*nz + c]
    double* const Sq = Sx + (size_t)NS * nz * BLOCK;           // Sq[q*nz + c]
    // Dc[k*nz + c] = d c_k / d z_c overwrites the sensitivity block column by column once the CQAs are known
    // (recourse_sens_rows() rows: the ng rows of Dc plus one spare row for the softmax-weighted gradient)
    double* const Dc = Sx;
    const double h = 1.0 / (double)img.nfe;
    const int NV = recourse_nv(nz);

    const long long t = blockIdx.x;
    const bool per_scen = ra.per_scenario == 1;
    long long i_d, j, grp;
    bool valid;
    if (per_scen) {
        // thread k of the launch owns the k-th running scenario
        const long long n_run = ra.n_active ? (long long)*ra.n_active : ra.n_d * ra.n_t;
        const long long k0 = t * (long long)BLOCK;
        if (k0 >= n_run) return;                                 // block-uniform
        valid = k0 + tid < n_run;
        const long long k = valid ? k0 + tid : n_run - 1;        // idle lanes replay the last scenario
        grp = ra.active ? ra.active[k] : k;
        i_d = grp / ra.n_t;
        j = grp % ra.n_t;
    } else if (ra.per_scenario == 2) {
        // one group over the whole grid: tiles run over the flattened (design point, theta) index, so a small theta
        // set (50 in the shipped scripts) does not leave most lanes of a tile idle
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
        if (ra.status[grp] != 0) return;                         // block-uniform: group already finished
    }
    const bool active = valid && ra.status[grp] == 0;
    if (!__syncthreads_or(active)) return;

    for (int c = 0; c < img.p_dim; ++c) S[tab_par[img.d_dim + c] * BLOCK] = __ldg(ra.theta + j * img.p_dim + c);
    for (int c = 0; c < img.d_dim; ++c) S[tab_par[c] * BLOCK] = __ldg(ra.d + i_d * img.d_dim + c);
    const double* zg = ra.z_try + grp * nz;

    run_ops(base, cz, img.seg[0], img.seg[1]);
    double xs[NS];
#pragma unroll
    for (int n = 0; n < NS; ++n) xs[n] = vm_ld(vm, tab_x0[n]);
    for (int q = 0; q < nq; ++q) S[(img.qe_slot + q) * BLOCK] = vm_ld(vm, tab_q0[q]);
    for (int i = 0; i < (NS + nq) * nz; ++i) Sx[(size_t)i * BLOCK] = 0.0;

    bool failed = false;
    unsigned my_iters = 0;
    for (int e = 0; e < img.nfe; ++e) {
        for (int c = 0; c < img.nce; ++c) {
            const int zi = tab_cmap[e * img.nce + c];
            S[(img.ctrl_slot + c) * BLOCK] = tab_zfix[zi] ? vm_ld(vm, tab_z[zi]) : __ldg(zg + zi);
        }
        run_ops(base, cz, img.seg[1], img.seg[2]);
        double xv[3][NS];
#pragma unroll
        for (int n = 0; n < NS; ++n) {
            xv[0][n] = xv[1][n] = xv[2][n] = xs[n];
            Xs[(3 * n + 0) * BLOCK] = xs[n];
            Xs[(3 * n + 1) * BLOCK] = xs[n];
            Xs[(3 * n + 2) * BLOCK] = xs[n];
        }
        bool conv = !active, stop = false;
        double s_prev = 0.0;
        double M[TRI ? 1 : N][TRI ? 1 : N], inv[TRI ? 1 : N], r[N];
        double T[TRI ? NS : 1][3][3], tinv[TRI ? NS : 1][3], Joff[TRI ? NS : 1][TRI ? NS : 1][3];
        // solve the Newton matrix (factors of the last assembly) for the right-hand side r, in place
        auto solve = [&](double (&rr)[N]) {
            if constexpr (TRI) {
#pragma unroll
                for (int n = 0; n < NS; ++n) {
                    double rn[3];
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj) {
                        rn[jj] = rr[jj * NS + n];
#pragma unroll
                        for (int m = 0; m < n; ++m) rn[jj] = fma(Joff[n][m][jj], rr[jj * NS + m], rn[jj]);
                    }
                    lu_apply<3>(T[n], tinv[n], rn);
#pragma unroll
                    for (int jj = 0; jj < 3; ++jj) rr[jj * NS + n] = rn[jj];
                }
            } else {
                lu_apply<N>(M, inv, rr);
            }
        };
        for (int it = 0;; ++it) {
            run_ops(base, cz, img.seg[2], img.seg[3]);
#pragma unroll
            for (int jj = 0; jj < 3; ++jj)
#pragma unroll
                for (int n = 0; n < NS; ++n) {
                    double lin = img.adot[jj + 1] * xs[n];
#pragma unroll
                    for (int kk = 0; kk < 3; ++kk) lin = fma(img.adot[4 * (kk + 1) + jj + 1], xv[kk][n], lin);
                    r[jj * NS + n] = fma(h, PT_LD(img.blk_f_o[0][n][jj]), -lin);
                    if constexpr (TRI) {
#pragma unroll
                        for (int kk = 0; kk < 3; ++kk) T[n][jj][kk] = img.adot[4 * (kk + 1) + jj + 1];
                        if ((img.blk_jmask[0] >> (n * NS + n)) & 1u)
                            T[n][jj][jj] = fma(-h, PT_LD(img.blk_j_o[0][n * NS + n][jj]), T[n][jj][jj]);
#pragma unroll
                        for (int m = 0; m < n; ++m)
                            Joff[n][m][jj] = ((img.blk_jmask[0] >> (n * NS + m)) & 1u) ? h * PT_LD(img.blk_j_o[0][n * NS + m][jj]) : 0.0;
                    } else {
#pragma unroll
                        for (int kk = 0; kk < 3; ++kk)
#pragma unroll
                            for (int m = 0; m < NS; ++m) {
                                const double a = (n == m) ? img.adot[4 * (kk + 1) + jj + 1] : 0.0;
                                M[jj * NS + n][kk * NS + m] = a;
                                if (kk == jj && ((img.blk_jmask[0] >> (n * NS + m)) & 1u))
                                    M[jj * NS + n][kk * NS + m] =
                                        fma(-h, PT_LD(img.blk_j_o[0][n * NS + m][jj]), a);
                            }
                    }
                }
            // at the last pass: factors at the converged point
            if constexpr (TRI) {
#pragma unroll
                for (int n = 0; n < NS; ++n) lu_factor<3>(T[n], tinv[n]);
            } else {
                lu_factor<N>(M, inv);
            }
            if (stop) break;
            solve(r);
            double s = 0.0, sc = 0.0;
#pragma unroll
            for (int kk = 0; kk < 3; ++kk)
#pragma unroll
                for (int n = 0; n < NS; ++n) {
                    const double dx = r[kk * NS + n];
                    const double xn = xv[kk][n] + dx;
                    xv[kk][n] = xn;
                    Xs[(3 * n + kk) * BLOCK] = xn;
                    s = fma(dx, dx, s);
                    sc = fma(xn, xn, sc);
                }
            if (!conv) ++my_iters;
            const double lim = img.rtol2 * fmax(sc, 1.0);
            const bool bad = !(s <= DBL_MAX);
            const bool ok = (s <= lim) || (s * s <= lim * s_prev && s <= 0.01 * s_prev);
            s_prev = s;
            if (bad && active) failed = true;
            if (ok || bad) conv = true;
            stop = __all_sync(0xffffffffu, conv) || (it + 1 >= img.max_it);
        }
        if (!conv) failed = true;

        // quadratures at the converged points
        for (int q = 0; q < nq; ++q) {
            if (!((img.fq_mask >> q) & 1u)) continue;
            double acc = img.bw[0] * PT_LD(img.fq_o[q][0]);
            acc = fma(img.bw[1], PT_LD(img.fq_o[q][1]), acc);
            acc = fma(img.bw[2], PT_LD(img.fq_o[q][2]), acc);
            S[(img.qe_slot + q) * BLOCK] = fma(h, acc, S[(img.qe_slot + q) * BLOCK]);
        }
        // forward sensitivities of this element with respect to every control that has acted so far.  The non-zero
        // entries of df/dz, dfq/dx and dfq/dz are compact lists in the image (sens_e): no mask or table walking here.
        const unsigned live = tab_live[e];
        int cact[4];                                            // z index of control function k on this element
#pragma unroll
        for (int k = 0; k < 4; ++k) cact[k] = k < img.nce ? (int)tab_cmap[e * img.nce + k] : -1;
        auto ctrl_of = [&](int k) { return k == 0 ? cact[0] : (k == 1 ? cact[1] : (k == 2 ? cact[2] : cact[3])); };
        // direct control terms of the quadratures (independent of the state sensitivities)
        for (int i = img.n_fz + img.n_fqx; i < img.n_fz + img.n_fqx + img.n_fqz; ++i) {
            const SensEntry en = img.sens_e[i];
            double acc = img.bw[0] * PT_LD(en.o[0]);
            acc = fma(img.bw[1], PT_LD(en.o[1]), acc);
            acc = fma(img.bw[2], PT_LD(en.o[2]), acc);
            double* sq = Sq + (size_t)(en.row * nz + ctrl_of(en.col)) * BLOCK;
            *sq = fma(h, acc, *sq);
        }
        for (int c = 0; c < nz; ++c) {
            if (!((live >> c) & 1u)) continue;
#pragma unroll
            for (int jj = 0; jj < 3; ++jj)
#pragma unroll
                for (int n = 0; n < NS; ++n) r[jj * NS + n] = -img.adot[jj + 1] * Sx[(size_t)(n * nz + c) * BLOCK];
            for (int i = 0; i < img.n_fz; ++i) {                 // h * df_n/dz_c where control c acts on this element
                const SensEntry en = img.sens_e[i];
                if (ctrl_of(en.col) != c) continue;
                const double f0 = h * PT_LD(en.o[0]), f1 = h * PT_LD(en.o[1]), f2 = h * PT_LD(en.o[2]);
#pragma unroll
                for (int n = 0; n < NS; ++n)
                    if (n == en.row) { r[n] += f0; r[NS + n] += f1; r[2 * NS + n] += f2; }
            }
            solve(r);
#pragma unroll
            for (int n = 0; n < NS; ++n) Sx[(size_t)(n * nz + c) * BLOCK] = r[2 * NS + n];
            for (int i = img.n_fz; i < img.n_fz + img.n_fqx; ++i) {      // quadratures: h * sum_j bw_j dfq/dx_n dX_n(j)
                const SensEntry en = img.sens_e[i];
                double d0 = 0.0, d1 = 0.0, d2 = 0.0;
#pragma unroll
                for (int n = 0; n < NS; ++n)
                    if (n == en.col) { d0 = r[n]; d1 = r[NS + n]; d2 = r[2 * NS + n]; }
                double acc = img.bw[0] * PT_LD(en.o[0]) * d0;
                acc = fma(img.bw[1] * PT_LD(en.o[1]), d1, acc);
                acc = fma(img.bw[2] * PT_LD(en.o[2]), d2, acc);
                double* sq = Sq + (size_t)(en.row * nz + c) * BLOCK;
                *sq = fma(h, acc, *sq);
            }
        }
#pragma unroll
        for (int n = 0; n < NS; ++n) xs[n] = xv[2][n];
    }

    // ---- scaled constraints c_k = g_k / M_k and their control gradients ----
#pragma unroll
    for (int n = 0; n < NS; ++n) S[(img.xe_slot + n) * BLOCK] = xs[n];
    run_ops(base, cz, img.seg[3], img.seg[4]);
    double cmax = 0.0;                       // max(0, max_k c_k)
    double ck[8];
    for (int k = 0; k < ng; ++k) {
        ck[k] = vm_ld(vm, tab_g[k]) * img.bigm_inv[k];
        cmax = fmax(cmax, ck[k]);
    }
    for (int c = 0; c < nz; ++c) {           // column c of Dc needs column c of [Sx; Sq] only: in place
        double dcv[8];
        for (int k = 0; k < ng; ++k) {
            double v = 0.0;
            if (!tab_zfix[c]) {
#pragma unroll
                for (int n = 0; n < NS; ++n) v = fma(vm_ld(vm, tab_gx[k * NS + n]), Sx[(size_t)(n * nz + c) * BLOCK], v);
                for (int q = 0; q < nq; ++q) v = fma(vm_ld(vm, tab_gq[k * nq + q]), Sq[(size_t)(q * nz + c) * BLOCK], v);
                v *= img.bigm_inv[k];
            }
            dcv[k] = v;
        }
        for (int k = 0; k < ng; ++k) Dc[(size_t)(k * nz + c) * BLOCK] = dcv[k];
    }
    recourse_emit(ra, nz, ng, NV, active, failed, ck, cmax, Dc, wsum, img.n_units < 32 ? img.n_units : 32, grp, j, t, my_iters);
}

struct LmArgs {
    long long n_groups;
    int nz, ng, n_tiles;       // n_tiles = partial sums per group (shared modes)
    int per_scenario;          // acc holds the group's own {failed, c_k, grad c_k} instead of reduced statistics
    int n_stages;              // shared modes: stages per evaluation in acc
    const double* acc;         // shared: (G, n_tiles, n_stages, NV) ; per-scenario: (G, recourse_nvp)
    const double* lo;
    const double* hi;
    const unsigned char* zfix;
    double* z_cur;
    double* z_try;
    double* F_cur;
    double* raw_cur;
    double* gr;                // (G, nz)
    double* H;                 // (G, nz, nz)
    double* lam;
    double* mu;
    int* status;
    int* first;
    int* iters;
    int* n_running;            // device counter of groups still running after this update
    // compaction (per-scenario mode): thread k updates group active_in[k], k < *n_active_in (null = identity over
    // all groups); groups that keep running are appended to active_out (warp-aggregated atomics on n_running)
    const long long* active_in;
    const int* n_active_in;
    long long* active_out;
    double mu_min, lam0, step_tol;
    int max_iter;
};

__device__ bool lm_update_group(const LmArgs& a, long long g);
template <int NZC, int NGC = 8>
__device__ bool lm_update_group_regs(const LmArgs& a, long long g, int* want_stage = nullptr, bool resumed = false);

// One thread per (running) group.  NZC > 0: the register-resident update for nz <= NZC controls (lm_update_group_regs);
// NZC = 0: the generic one (local-memory arrays, any nz <= MAX_NZ).
// NGC: compile-time bound on the constraints per scenario (per-scenario mode keeps c_k and their gradients in registers).
template <int NZC, int NGC = 8>
#ifndef PYDSB_LM_CTAS
#define PYDSB_LM_CTAS 2
#endif
__global__ void __launch_bounds__(128, NZC > 0 ? PYDSB_LM_CTAS : 1) lm_update_kernel(const LmArgs a)
{
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long n_in = a.n_active_in ? (long long)*a.n_active_in : a.n_groups;
    bool done = true;                          // lanes without a running group append nothing
    long long g = 0;
    if (k < n_in) {
        g = a.active_in ? a.active_in[k] : k;
        if (a.status[g] == 0) {
            if constexpr (NZC > 0) done = lm_update_group_regs<NZC, NGC>(a, g);
            else done = lm_update_group(a, g);
        }
    }
    // running groups: count them, and (per-scenario mode) append them to the next active list
    const unsigned run_mask = __ballot_sync(0xffffffffu, !done);
    if (run_mask) {
        const int lane = threadIdx.x & 31;
        const int leader = __ffs(run_mask) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(a.n_running, __popc(run_mask));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (!done && a.active_out) a.active_out[base + __popc(run_mask & ((1u << lane) - 1u))] = g;
    }
}

// Statistics of one group at continuation stage `stage` (mu = the stage's value): {Phi_mu, Phi, grad} here, the
// Gauss-Newton H on demand (lm_group_hessian) -- it is only needed when a step is computed on a non-empty free set,
// and the bang-bang optima of the shipped problems end every stage with all controls on their bounds.
__device__ void lm_group_stats(const LmArgs& a, long long g, int stage, double mu, double& F_t, double& raw_t, double* g_t,
                               double* pk)
{
    const int nz = a.nz;
    F_t = 0.0;
    raw_t = 0.0;
    for (int c = 0; c < nz; ++c) g_t[c] = 0.0;
    if (a.per_scenario) {
        // the scenario's own c_k and grad c_k (weight 1): same formulas as the epilogue of sens_kernel
        const int ng = a.ng;
        const double* p = a.acc + g * recourse_nvp(ng, nz);
        const bool failed = p[0] != 0.0;
        const double* ck = p + 1;
        const double* Dc = p + 1 + ng;
        double cmax = 0.0;
        for (int k = 0; k < ng; ++k) cmax = fmax(cmax, ck[k]);
        hinge_point(ng, ck, cmax, failed, mu, pk, F_t, raw_t);
        for (int c = 0; c < nz; ++c)
            for (int k = 0; k < ng; ++k) g_t[c] = fma(pk[k], Dc[k * nz + c], g_t[c]);
    } else {
        const int NV = recourse_nv(nz);
        for (int t = 0; t < a.n_tiles; ++t) {
            const double* p = a.acc + ((g * a.n_tiles + t) * a.n_stages + stage) * NV;
            F_t += p[0];
            raw_t += p[1];
            for (int c = 0; c < nz; ++c) g_t[c] += p[2 + c];
        }
    }
}

__device__ void lm_group_hessian(const LmArgs& a, long long g, int stage, double mu, const double* g_t, const double* pk,
                                 double* H_t)
{
    const int nz = a.nz;
    for (int i = 0; i < nz * nz; ++i) H_t[i] = 0.0;
    if (a.per_scenario) {
        const int ng = a.ng;
        const double* Dc = a.acc + g * recourse_nvp(ng, nz) + 1 + ng;
        for (int r = 0; r < nz; ++r)
            for (int c = r; c < nz; ++c) {
                double hv = 0.0;
                for (int k = 0; k < ng; ++k) hv = fma(pk[k] * Dc[k * nz + r], Dc[k * nz + c], hv);
                H_t[r * nz + c] = fma(-g_t[r], g_t[c], hv) / mu;
            }
    } else {
        const int NV = recourse_nv(nz);
        for (int t = 0; t < a.n_tiles; ++t) {
            const double* p = a.acc + ((g * a.n_tiles + t) * a.n_stages + stage) * NV;
            int idx = 2 + nz;
            for (int r = 0; r < nz; ++r)
                for (int c = r; c < nz; ++c, ++idx) H_t[r * nz + c] += p[idx];
        }
    }
    for (int r = 0; r < nz; ++r)
        for (int c = 0; c < r; ++c) H_t[r * nz + c] = H_t[c * nz + r];
}

// Per-scenario mode: is the softmax over {0, c_1..c_K} / mu one-hot to far below round-off (every other term < 1e-30
// of the largest)?  Then it stays so for every smaller mu.
__device__ bool lm_softmax_saturated(const LmArgs& a, long long g, double mu)
{
    const int ng = a.ng;
    const double* p = a.acc + g * recourse_nvp(ng, a.nz);
    if (p[0] != 0.0) return false;                        // failed state solve
    const double* ck = p + 1;
    double cmax = 0.0;
    int kmax = -1;
    for (int k = 0; k < ng; ++k)
        if (ck[k] > cmax) { cmax = ck[k]; kmax = k; }
    if (kmax < 0) return false;
    double rest = exp(-cmax / mu);
    for (int k = 0; k < ng; ++k)
        if (k != kmax) rest += exp((ck[k] - cmax) / mu);
    return rest < 1e-30;
}

// The accept/reject + step logic of oracle/recourse.py::minimize_hinge for one group; true when it finished.
// One call consumes one evaluation (state + sensitivity solve at z_try).  When the group converges at the current
// mu with z_cur = the evaluated point, the next stage of the continuation needs the statistics at the same controls
// for the smaller mu -- they are in hand (see RECOURSE_STAGES), so the stage loop continues without a new evaluation.
__device__ bool lm_update_group(const LmArgs& a, long long g)
{
    const int nz = a.nz;
    double* zc = a.z_cur + g * nz;
    double* zt = a.z_try + g * nz;
    double* gr = a.gr + g * nz;
    double* H = a.H + g * nz * nz;
    a.iters[g] += 1;
    bool done = false;
    const int stages_in_hand = a.per_scenario ? (1 << 30) : a.n_stages;
    for (int stage = 0;; ++stage) {
    double F_t, raw_t, g_t[MAX_NZ], pk[8];
    const double mu_stage = a.mu[g];
    lm_group_stats(a, g, stage, mu_stage, F_t, raw_t, g_t, pk);
    bool converged_here = false;
    bool at_eval_point = false;               // z_cur is the point the statistics were evaluated at
    bool pinned = false;                      // converged with every control on a bound (empty free set)
    if (raw_t == 0.0) {                       // every scenario of the group is feasible: Phi = 0 exactly
        for (int c = 0; c < nz; ++c) zc[c] = zt[c];
        a.raw_cur[g] = 0.0;
        a.status[g] = 1;
        done = true;
    }
    if (!done) {
        const bool first = a.first[g] != 0;
        if (first || F_t < a.F_cur[g]) {
            double moved = 0.0;
            for (int c = 0; c < nz; ++c) {
                const double span = (a.hi[c] - a.lo[c] <= DBL_MAX) ? (a.hi[c] - a.lo[c]) : 1.0;
                moved = fmax(moved, fabs(zt[c] - zc[c]) / span);
            }
            const bool small = !first && (moved <= a.step_tol || a.F_cur[g] - F_t <= 1e-13 * fmax(F_t, 1e-300));
            for (int c = 0; c < nz; ++c) { zc[c] = zt[c]; gr[c] = g_t[c]; }
            a.F_cur[g] = F_t;
            a.raw_cur[g] = raw_t;
            if (!first) a.lam[g] = fmax(a.lam[g] / 3.0, 1e-12);
            a.first[g] = 0;
            converged_here = small;
            at_eval_point = true;
        } else {
            a.lam[g] = a.lam[g] * 4.0;
            if (a.lam[g] >= 1e12) converged_here = true;
        }
        if (!converged_here) {
            // projected LM step on the free set
            int idx[MAX_NZ], nf = 0;
            for (int c = 0; c < nz; ++c) {
                const bool at_lo = zc[c] <= a.lo[c] && gr[c] > 0.0;
                const bool at_hi = zc[c] >= a.hi[c] && gr[c] < 0.0;
                if (!a.zfix[c] && !at_lo && !at_hi) idx[nf++] = c;
            }
            double step[MAX_NZ];
            for (int c = 0; c < nz; ++c) step[c] = 0.0;
            if (nf > 0) {
                if (at_eval_point) {                       // z_cur was just accepted: its H is not in global memory yet
                    double H_t[MAX_NZ * MAX_NZ];
                    lm_group_hessian(a, g, stage, mu_stage, g_t, pk, H_t);
                    for (int i = 0; i < nz * nz; ++i) H[i] = H_t[i];
                }
                // H is a weighted covariance of the constraint gradients divided by mu: positive semi-definite in exact
                // arithmetic.  Where one constraint dominates every scenario it is zero up to round-off of either
                // sign (~1e-16 gr_i^2 / mu): treat a diagonal entry below 1e-10 gr_i^2 / mu as "no curvature", drop the
                // rows / columns of those directions, so that the damped matrix stays positive definite and the step
                // along them goes to the bound.
                double hd[MAX_NZ], dmax = 0.0;
                for (int i = 0; i < nf; ++i) {
                    const double hii = H[idx[i] * nz + idx[i]], gi = gr[idx[i]];
                    hd[i] = hii > 1e-10 * gi * gi / a.mu[g] ? hii : 0.0;
                    dmax = fmax(dmax, hd[i]);
                }
                const double floor_ = 1e-8 * dmax + 1e-300;
                double l = a.lam[g];
                for (int attempt = 0; attempt < 12; ++attempt) {
                    double A[MAX_NZ * MAX_NZ];
                    for (int i = 0; i < nf; ++i)
                        for (int jx = 0; jx < nf; ++jx)
                            A[i * nf + jx] = (hd[i] > 0.0 && hd[jx] > 0.0) ? H[idx[i] * nz + idx[jx]] : 0.0;
                    for (int i = 0; i < nf; ++i) A[i * nf + i] = hd[i] + l * (hd[i] + floor_) + 1e-30;
                    bool okc = true;                       // in-place Cholesky A = L L^T
                    for (int i = 0; i < nf && okc; ++i)
                        for (int jx = 0; jx <= i; ++jx) {
                            double sum = A[i * nf + jx];
                            for (int k = 0; k < jx; ++k) sum -= A[i * nf + k] * A[jx * nf + k];
                            if (i == jx) {
                                if (!(sum > 0.0)) { okc = false; break; }
                                A[i * nf + i] = sqrt(sum);
                            } else {
                                A[i * nf + jx] = sum / A[jx * nf + jx];
                            }
                        }
                    if (!okc) { l *= 10.0; continue; }
                    double y[MAX_NZ];
                    for (int i = 0; i < nf; ++i) {
                        double sum = -gr[idx[i]];
                        for (int k = 0; k < i; ++k) sum -= A[i * nf + k] * y[k];
                        y[i] = sum / A[i * nf + i];
                    }
                    for (int i = nf - 1; i >= 0; --i) {
                        double sum = y[i];
                        for (int k = i + 1; k < nf; ++k) sum -= A[k * nf + i] * y[k];
                        y[i] = sum / A[i * nf + i];
                    }
                    for (int i = 0; i < nf; ++i) step[idx[i]] = y[i];
                    break;
                }
            }
            bool same = true;
            for (int c = 0; c < nz; ++c) {
                const double zn = fmin(fmax(zc[c] + step[c], a.lo[c]), a.hi[c]);
                if (zn != zc[c]) same = false;
                step[c] = zn;
            }
            if (same) { converged_here = true; pinned = nf == 0; }
            else for (int c = 0; c < nz; ++c) zt[c] = step[c];
        }
        if (converged_here) {
            if (a.mu[g] <= a.mu_min * 1.0001) {
                a.status[g] = 2;
                done = true;
            } else if (a.per_scenario && pinned && at_eval_point && lm_softmax_saturated(a, g, a.mu[g])) {
                // bang-bang scenario whose softmax is one-hot already: a smaller mu changes neither the gradient nor
                // the (empty) free set, every remaining stage would converge on the spot
                a.mu[g] = a.mu_min;
                a.status[g] = 2;
                done = true;
            } else {
                a.mu[g] = fmax(a.mu[g] / 10.0, a.mu_min);
                a.first[g] = 1;
                a.lam[g] = a.lam0;
                for (int c = 0; c < nz; ++c) zt[c] = zc[c];
                // the next stage evaluates z_try = z_cur at the smaller mu: in hand if that is the evaluated point
                if (at_eval_point && stage + 1 < stages_in_hand) continue;
            }
        }
        if (!done && a.iters[g] >= a.max_iter) {
            a.status[g] = 3;
            done = true;
        }
    }
    break;
    }
    return done;
}

// lm_update_group with the group's whole state in registers (nz <= NZC): every per-group value is loaded up front
// (independent loads in flight together instead of a chain of dependent global accesses ordered by possible aliasing),
// all arrays are statically indexed (the free set is a bit mask, the damped system is the full nz x nz matrix with
// identity rows for the controls that are not free -- the same factors and the same step as the compacted system),
// and the state is written back once at the end.  Same decisions and, up to FMA contraction, the same numbers.
// want_stage (rfused_kernel): the caller emits ONE stage per call; when the group converges at the evaluated point and the
// continuation goes on, *want_stage is set instead of reading the next stage -- the caller emits the statistics for the new
// mu at the same point (nothing but the smoothing changes) and calls again with resumed = true, which is the `continue` of
// the stage loop below: the same numbers, the same decisions, no iteration counted in between.
template <int NZC, int NGC>
__device__ bool lm_update_group_regs(const LmArgs& a, long long g, int* want_stage, bool resumed)
{
    const int nz = a.nz, ng = a.ng;
    const double* __restrict__ acc = a.acc;
    double zc[NZC], zt[NZC], gr[NZC];
    // bounds: uniform across the threads, read where they are used (L1-resident) instead of held in 32 registers
    auto lo = [&](int c) { return c < nz ? __ldg(a.lo + c) : 0.0; };
    auto hi = [&](int c) { return c < nz ? __ldg(a.hi + c) : 0.0; };
    bool fixd[NZC];
#pragma unroll
    for (int c = 0; c < NZC; ++c) {
        const bool in = c < nz;
        zc[c] = in ? a.z_cur[g * nz + c] : 0.0;
        zt[c] = in ? a.z_try[g * nz + c] : 0.0;
        gr[c] = in ? a.gr[g * nz + c] : 0.0;
        fixd[c] = in ? a.zfix[c] != 0 : true;
    }
    const int iters = a.iters[g] + (resumed ? 0 : 1);
    int first = a.first[g], status = 0;
    bool lazy = false;
    double F_cur = a.F_cur[g], raw_cur = a.raw_cur[g], lam = a.lam[g], mu = a.mu[g];
    double* const H = a.H + g * nz * nz;
    // per-scenario mode: the scenario's own record
    const double* const rec = acc + g * recourse_nvp(ng, nz);
    bool rec_failed = false;
    double ck[NGC], cmax = 0.0;
    // per-scenario mode: the scenario's gradients d c_k / d z_c -- loaded once when they fit the registers (few constraints)
    constexpr bool PRE = NGC <= 2;
    double Dr[PRE ? NGC : 1][PRE ? NZC : 1];
    const double* const Dc = rec + 1 + ng;
    if (a.per_scenario) {
        rec_failed = rec[0] != 0.0;
#pragma unroll
        for (int k = 0; k < NGC; ++k) {
            ck[k] = k < ng ? rec[1 + k] : 0.0;
            if (k < ng) cmax = fmax(cmax, ck[k]);
            if constexpr (PRE) {
#pragma unroll
                for (int c = 0; c < NZC; ++c) Dr[k][c] = (k < ng && c < nz) ? Dc[k * nz + c] : 0.0;
            }
        }
    }
    auto dval = [&](int k, int c) -> double {
        if constexpr (PRE) return Dr[k][c];
        else return Dc[k * nz + c];
    };
    const int NV = recourse_nv(nz);

    bool done = false;
    const int stages_in_hand = a.per_scenario ? (1 << 30) : a.n_stages;
    for (int stage = 0;; ++stage) {
        // ---- statistics of this stage: {Phi_mu, Phi, grad} ----
        double F_t = 0.0, raw_t = 0.0, g_t[NZC], pk[NGC];
        double e0 = 0.0, ek[PRE ? NGC : 1];         // the unnormalised softmax terms of this stage (saturation test below)
#pragma unroll
        for (int c = 0; c < NZC; ++c) g_t[c] = 0.0;
        const double mu_stage = mu;
        if (a.per_scenario) {
            if (rec_failed) {
                F_t = raw_t = 1.0;
#pragma unroll
                for (int k = 0; k < NGC; ++k) pk[k] = 0.0;
            } else {
                e0 = exp(-cmax / mu_stage);
                double Z = e0;
#pragma unroll
                for (int k = 0; k < NGC; ++k) {
                    // (the largest c_k, when positive, is cmax itself: exp(0) needs no call)
                    pk[k] = k < ng ? (ck[k] == cmax ? 1.0 : exp((ck[k] - cmax) / mu_stage)) : 0.0;
                    if (k < ng) Z += pk[k];
                    if constexpr (PRE) ek[k] = pk[k];
                }
                const double zi = 1.0 / Z;
#pragma unroll
                for (int k = 0; k < NGC; ++k) pk[k] *= zi;
                F_t = cmax + mu_stage * log(Z);
                raw_t = cmax;
            }
#pragma unroll
            for (int c = 0; c < NZC; ++c)
                if (c < nz) {
#pragma unroll
                    for (int k = 0; k < NGC; ++k)
                        if (k < ng) g_t[c] = fma(pk[k], dval(k, c), g_t[c]);
                }
        } else {
            for (int t = 0; t < a.n_tiles; ++t) {
                const double* p = acc + ((g * a.n_tiles + t) * a.n_stages + stage) * NV;
                F_t += p[0];
                raw_t += p[1];
#pragma unroll
                for (int c = 0; c < NZC; ++c)
                    if (c < nz) g_t[c] += p[2 + c];
            }
        }
        bool converged_here = false, at_eval_point = false, pinned = false;
        if (raw_t == 0.0) {                       // every scenario of the group is feasible: Phi = 0 exactly
#pragma unroll
            for (int c = 0; c < NZC; ++c) zc[c] = zt[c];
            raw_cur = 0.0;
            status = 1;
            done = true;
        }
        if (!done) {
            if (first || F_t < F_cur) {
                double moved = 0.0;
#pragma unroll
                for (int c = 0; c < NZC; ++c)
                    if (c < nz) {
                        const double span = (hi(c) - lo(c) <= DBL_MAX) ? (hi(c) - lo(c)) : 1.0;
                        moved = fmax(moved, fabs(zt[c] - zc[c]) / span);
                    }
                const bool small = !first && (moved <= a.step_tol || F_cur - F_t <= 1e-13 * fmax(F_t, 1e-300));
#pragma unroll
                for (int c = 0; c < NZC; ++c) { zc[c] = zt[c]; gr[c] = g_t[c]; }
                F_cur = F_t;
                raw_cur = raw_t;
                if (!first) lam = fmax(lam / 3.0, 1e-12);
                first = 0;
                converged_here = small;
                at_eval_point = true;
            } else {
                lam = lam * 4.0;
                if (lam >= 1e12) converged_here = true;
            }
            if (!converged_here) {
                // projected LM step on the free set
                unsigned fm = 0;
#pragma unroll
                for (int c = 0; c < NZC; ++c) {
                    const bool at_lo = zc[c] <= lo(c) && gr[c] > 0.0;
                    const bool at_hi = zc[c] >= hi(c) && gr[c] < 0.0;
                    if (!fixd[c] && !at_lo && !at_hi) fm |= 1u << c;
                }
                double step[NZC];
#pragma unroll
                for (int c = 0; c < NZC; ++c) step[c] = 0.0;
                if (fm) {
                    if (at_eval_point) {                   // z_cur was just accepted: its H is not in global memory yet
                        if (a.per_scenario) {
                            const double inv_mu = 1.0 / mu_stage;
#pragma unroll
                            for (int r = 0; r < NZC; ++r)
#pragma unroll
                                for (int c = r; c < NZC; ++c)
                                    if (((fm >> r) & (fm >> c)) & 1u) {    // only free x free entries are ever read,
                                        double hv = 0.0;                   // and only from the lower triangle
#pragma unroll
                                        for (int k = 0; k < NGC; ++k)
                                            if (k < ng) hv = fma(pk[k] * dval(k, r), dval(k, c), hv);
                                        H[c * nz + r] = fma(-g_t[r], g_t[c], hv) * inv_mu;
                                    }
                        } else {
#pragma unroll
                            for (int r = 0; r < NZC; ++r)
#pragma unroll
                                for (int c = r; c < NZC; ++c)
                                    if (((fm >> r) & (fm >> c)) & 1u) {
                                        const int idx = 2 + nz + r * nz - (r * (r - 1)) / 2 + (c - r);   // packed upper triangle
                                        double s = 0.0;
                                        for (int t = 0; t < a.n_tiles; ++t)
                                            s += acc[((g * a.n_tiles + t) * a.n_stages + stage) * NV + idx];
                                        H[c * nz + r] = s;
                                    }
                        }
                    }
                    // see lm_update_group for the "no curvature" rule on the diagonal
                    double hd[NZC], dmax = 0.0;
#pragma unroll
                    for (int c = 0; c < NZC; ++c) {
                        hd[c] = 0.0;
                        if ((fm >> c) & 1u) {
                            const double hii = H[c * nz + c];
                            hd[c] = hii > 1e-10 * gr[c] * gr[c] / mu ? hii : 0.0;
                            dmax = fmax(dmax, hd[c]);
                        }
                    }
                    const double floor_ = 1e-8 * dmax + 1e-300;
                    double l = lam;
#pragma unroll 1                                             /* (the retries are rare: one copy of the factorisation) */
                    for (int attempt = 0; attempt < 12; ++attempt) {
                        double A[NZC][NZC];                // lower triangle used
#pragma unroll
                        for (int i = 0; i < NZC; ++i)
#pragma unroll
                            for (int jx = 0; jx <= i; ++jx) {
                                const bool fi = (fm >> i) & 1u, fj = (fm >> jx) & 1u;
                                double v = 0.0;
                                if (i == jx) v = fi ? hd[i] + l * (hd[i] + floor_) + 1e-30 : 1.0;
                                else if (fi && fj && hd[i] > 0.0 && hd[jx] > 0.0) v = H[i * nz + jx];
                                A[i][jx] = v;
                            }
                        // in-place Cholesky A = L L^T; the diagonal holds 1 / L_ii (one rsqrt per row instead of a
                        // square root and a division per entry)
                        bool okc = true;
#pragma unroll
                        for (int i = 0; i < NZC; ++i)
#pragma unroll
                            for (int jx = 0; jx <= i; ++jx) {
                                double sum = A[i][jx];
#pragma unroll
                                for (int k = 0; k < jx; ++k) sum -= A[i][k] * A[jx][k];
                                if (i == jx) {
                                    if (!(sum > 0.0)) okc = false;
                                    A[i][i] = rsqrt(sum);
                                } else {
                                    A[i][jx] = sum * A[jx][jx];
                                }
                            }
                        if (!okc) { l *= 10.0; continue; }
                        double y[NZC];
#pragma unroll
                        for (int i = 0; i < NZC; ++i) {
                            double sum = ((fm >> i) & 1u) ? -gr[i] : 0.0;
#pragma unroll
                            for (int k = 0; k < i; ++k) sum -= A[i][k] * y[k];
                            y[i] = sum * A[i][i];
                        }
#pragma unroll
                        for (int i = NZC - 1; i >= 0; --i) {
                            double sum = y[i];
#pragma unroll
                            for (int k = i + 1; k < NZC; ++k) sum -= A[k][i] * y[k];
                            y[i] = sum * A[i][i];
                        }
#pragma unroll
                        for (int c = 0; c < NZC; ++c) step[c] = ((fm >> c) & 1u) ? y[c] : 0.0;
                        break;
                    }
                }
                bool same = true;
#pragma unroll
                for (int c = 0; c < NZC; ++c) {
                    const double zn = fmin(fmax(zc[c] + step[c], lo(c)), hi(c));
                    if (zn != zc[c]) same = false;
                    step[c] = zn;
                }
                if (same) { converged_here = true; pinned = fm == 0; }
                else {
#pragma unroll
                    for (int c = 0; c < NZC; ++c) zt[c] = step[c];
                }
            }
            if (converged_here) {
                bool saturated = false;
                if (a.per_scenario && pinned && at_eval_point && !rec_failed && mu > a.mu_min * 1.0001) {
                    // lm_softmax_saturated on the values in hand
                    int kmax = -1;
                    double cm = 0.0;
#pragma unroll
                    for (int k = 0; k < NGC; ++k)
                        if (k < ng && ck[k] > cm) { cm = ck[k]; kmax = k; }
                    if (kmax >= 0) {
                        // cm = cmax and mu = this stage's: the terms are the ones the statistics were formed from
                        double rest = PRE ? e0 : exp(-cm / mu);
#pragma unroll
                        for (int k = 0; k < NGC; ++k)
                            if (k < ng && k != kmax) {
                                if constexpr (PRE) rest += ek[k];
                                else rest += exp((ck[k] - cm) / mu);
                            }
                        saturated = rest < 1e-30;
                    }
                }
                if (mu <= a.mu_min * 1.0001) {
                    status = 2;
                    done = true;
                } else if (saturated) {
                    mu = a.mu_min;
                    status = 2;
                    done = true;
                } else {
                    mu = fmax(mu / 10.0, a.mu_min);
                    first = 1;
                    lam = a.lam0;
#pragma unroll
                    for (int c = 0; c < NZC; ++c) zt[c] = zc[c];
                    if (at_eval_point && stage + 1 < stages_in_hand) continue;
                    if (at_eval_point && want_stage) { *want_stage = 1; lazy = true; }
                }
            }
            if (!done && !lazy && iters >= a.max_iter) {
                status = 3;
                done = true;
            }
        }
        break;
    }
    // ---- write the group's state back ----
#pragma unroll
    for (int c = 0; c < NZC; ++c)
        if (c < nz) {
            a.z_cur[g * nz + c] = zc[c];
            a.z_try[g * nz + c] = zt[c];
            a.gr[g * nz + c] = gr[c];
        }
    a.iters[g] = iters;
    a.first[g] = first;
    a.status[g] = status;
    a.F_cur[g] = F_cur;
    a.raw_cur[g] = raw_cur;
    a.lam[g] = lam;
    a.mu[g] = mu;
    return done;
}

// Sum the tile partials of the (single) global group: out[v] = sum_t acc[t * snv + v].  One CTA per value, a strided
// pass then a fixed-shape tree, so the result does not depend on scheduling.  (The per-group thread of
// lm_update_kernel would otherwise walk n_d * n_tiles * stages * NV values alone.)
__global__ void sum_tiles_kernel(const double* __restrict__ acc, int n_tiles, int snv, double* __restrict__ out)
{
    __shared__ double sh[128];
    const int v = blockIdx.x;
    double s = 0.0;
    for (int t = threadIdx.x; t < n_tiles; t += 128) s += acc[(size_t)t * snv + v];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 64; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[v] = sh[0];
}

__global__ void lm_init_kernel(long long n_groups, int nz, const double* z0, int z0_rows, double* z_cur, double* z_try,
                               double* F_cur, double* raw_cur, double* lam, double* mu, int* status, int* first, int* iters,
                               double lam0, double mu0)
{
    const long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    for (int c = 0; c < nz; ++c) {
        const double v = z0 ? z0[(z0_rows > 1 ? g : 0) * nz + c] : 0.0;
        z_cur[g * nz + c] = v;
        z_try[g * nz + c] = v;
    }
    F_cur[g] = DBL_MAX;
    raw_cur[g] = DBL_MAX;
    lam[g] = lam0;
    mu[g] = mu0;
    status[g] = 0;
    first[g] = 1;
    iters[g] = 0;
}

#undef PT_LD

}  // namespace pydsb
