// Feasibility kernel: collocation state solve + CQAs + big-M indicator + P(feasible) partials.
//
// Work decomposition (B200: 148 SMs, persistent CTAs):
//   tile  = one design point d_i  x  BLOCK consecutive theta samples
//   CTA   = BLOCK threads, loops over tiles  t = blockIdx.x, +gridDim.x, ...
//   thread = one (d_i, theta_j) scenario for the whole pipeline
//
// Per scenario (what the reference does inside one GAMS/CONOPT solve per design point,
// pyds/run.py:55-67 + pyds/solver.py:50-55, for fixed controls):
//   1. `pro` tape: parameter-only sub-expressions (Arrhenius factors etc.), initial states
//   2. for each finite element: `elem` tape, then full Newton on the 3*NS collocation unknowns:
//        `pt` tape (f, df/dx at the 3 Radau points, evaluated 3-wide) -> residual and Jacobian
//        assembled in registers -> register-resident dense LU -> update
//      quadrature states advance with the Radau weights
//   3. `g` tape -> constraint values; indicator = any(g_k > 0)   (utils.py:23-29, solver.py:71-80)
//   4. w_j * 1[feasible] reduced with warp shuffles; one partial per tile (deterministic finish
//      in reduce_partials_kernel) -- DEUS's efp reduction.
//
// theta tiles are staged by 1-D bulk TMA (cp.async.bulk + mbarrier), double buffered: the copy
// for tile k+1 is in flight while tile k is computed.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "tape_vm.cuh"

namespace pydsb {

struct ExecImage {
    // global-memory copies made by pydsb_model_create
    const uint64_t* code;      // pro | elem | pt | g
    const double* consts;      // constant pool
    const uint16_t* tabs;      // x0[ns] q0[nq] ztab[nz] zfix[nz] gtab[ng] cmap[nfe*nce]
    int n_pro, n_elem, n_pt, n_g, n_const, n_tabs;
    int ns, nq, ng, d_dim, p_dim, nz, nfe, nce, n_units;
    int param_slot, ctrl_slot, xe_slot, qe_slot, x_slot, f_slot, j_slot, fq_slot;
    unsigned j_mask;
    int fq_state_dep;
    int max_it;
    double rtol;
    double adot[16];
    double bw[3];
    // shared-memory carve-up (bytes from the start of dynamic smem)
    int off_code, off_const, off_tabs, off_land, off_slots, smem_bytes;
};

struct LaunchArgs {
    const double* d;
    const double* theta;
    const double* w;           // may be null -> uniform 1/n_t
    const double* z;
    int z_mode;
    long long n_d, n_t;
    int n_tiles;               // theta tiles per design point
    int theta_tma_ok;          // base pointer 16-byte aligned
    double* partial;           // (n_d, n_tiles)
    double* g_out;
    double* ind_out;
    unsigned long long* counters;
};

// ---------------------------------------------------------------------------------------------
// mbarrier / bulk-TMA helpers (sm_90+ PTX; SASS: SYNCS.*, UBLKCP)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    unsigned done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

// ---------------------------------------------------------------------------------------------
// Register-resident dense LU solve without pivoting (N = 3, 6, 9).  The Newton matrix
//   A^T (x) I - h * blockdiag(df/dx)
// has the fixed, well-conditioned collocation part on every diagonal block; a vanishing pivot
// produces inf/nan which the caller reports as a failed scenario.
// ---------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void lu_solve_inplace(double (&M)[N][N], double (&r)[N])
{
    double inv[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        inv[k] = fast_rcp(M[k][k]);
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const double l = M[i][k] * inv[k];
#pragma unroll
            for (int c = k + 1; c < N; ++c) M[i][c] = fma(-l, M[k][c], M[i][c]);
            r[i] = fma(-l, r[k], r[i]);
        }
    }
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
        double s = r[k];
#pragma unroll
        for (int c = k + 1; c < N; ++c) s = fma(-M[k][c], r[c], s);
        r[k] = s * inv[k];
    }
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned warp_sum_u32(unsigned v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NS>
__global__ void __launch_bounds__(BLOCK, (NS <= 2 ? 4 : 2))
    feas_kernel(const __grid_constant__ ExecImage img, const __grid_constant__ LaunchArgs la)
{
    constexpr int N = 3 * NS;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);                 // [2]
    double* red = reinterpret_cast<double*>(smem + 16);                 // [BLOCK/32]
    uint64_t* code = reinterpret_cast<uint64_t*>(smem + img.off_code);
    double* cpool = reinterpret_cast<double*>(smem + img.off_const) + 1;   // cpool[-1] = 0.0
    uint16_t* tabs = reinterpret_cast<uint16_t*>(smem + img.off_tabs);
    double* land = reinterpret_cast<double*>(smem + img.off_land);      // [2][BLOCK*p_dim]
    double* S = reinterpret_cast<double*>(smem + img.off_slots) + tid;

    // ---- stage the tapes, constants and tables in shared memory (once per CTA) ----
    const int n_code = img.n_pro + img.n_elem + img.n_pt + img.n_g;
    for (int i = tid; i < n_code; i += BLOCK) code[i] = img.code[i];
    for (int i = tid; i < img.n_const; i += BLOCK) cpool[i] = img.consts[i];
    for (int i = tid; i < img.n_tabs; i += BLOCK) tabs[i] = img.tabs[i];
    if (tid == 0) {
        cpool[-1] = 0.0;
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const uint64_t* code_pro = code;
    const uint64_t* code_elem = code_pro + img.n_pro;
    const uint64_t* code_pt = code_elem + img.n_elem;
    const uint64_t* code_g = code_pt + img.n_pt;
    const uint16_t* tab_x0 = tabs;
    const uint16_t* tab_q0 = tab_x0 + img.ns;
    const uint16_t* tab_z = tab_q0 + img.nq;
    const uint16_t* tab_zfix = tab_z + img.nz;
    const uint16_t* tab_g = tab_zfix + img.nz;
    const uint16_t* tab_cmap = tab_g + img.ng;

    VM vm;
    vm.S = S;
    vm.cpool = cpool;

    double* const Xs = S + img.x_slot * BLOCK;
    const double* const Fs = S + img.f_slot * BLOCK;
    const double* const Js = S + img.j_slot * BLOCK;
    const double* const FQs = S + img.fq_slot * BLOCK;
    double* const Ps = S + img.param_slot * BLOCK;

    const double h = 1.0 / (double)img.nfe;
    const int p_dim = img.p_dim, d_dim = img.d_dim;
    const unsigned tile_bytes_full = (unsigned)(BLOCK * p_dim * sizeof(double));

    const long long total = la.n_d * (long long)la.n_tiles;
    unsigned phase_bits = 0;

    auto tile_rows = [&](long long t) -> int {
        const long long j0 = (t % la.n_tiles) * (long long)BLOCK;
        const long long rem = la.n_t - j0;
        return rem < BLOCK ? (int)rem : BLOCK;
    };
    auto tile_tma = [&](int rows) -> bool {
        return la.theta_tma_ok && ((rows * p_dim * (int)sizeof(double)) % 16 == 0);
    };
    auto issue = [&](long long t, int b) {
        const int rows = tile_rows(t);
        if (tile_tma(rows)) {
            const long long j0 = (t % la.n_tiles) * (long long)BLOCK;
            const unsigned bytes = (unsigned)(rows * p_dim * sizeof(double));
            mbar_expect_tx(&mbar[b], bytes);
            tma_load_1d(land + (size_t)b * BLOCK * p_dim, la.theta + j0 * p_dim, bytes, &mbar[b]);
        }
    };

    long long t = blockIdx.x;
    if (t < total && tid == 0) issue(t, 0);

    unsigned long long cnt_iters = 0, cnt_warp_iters = 0;
    unsigned cnt_failed = 0, cnt_done = 0;

    for (int k = 0; t < total; t += gridDim.x, ++k) {
        const int b = k & 1;
        const long long i_d = t / la.n_tiles;
        const long long j0 = (t % la.n_tiles) * (long long)BLOCK;
        const int rows = tile_rows(t);
        const bool valid = tid < rows;
        const int row = valid ? tid : rows - 1;           // clamp: idle lanes replay the last row
        const long long j = j0 + row;

        // ---- parameters into the slot file ----
        if (tile_tma(rows)) {
            mbar_wait(&mbar[b], (phase_bits >> b) & 1u);
            phase_bits ^= 1u << b;
            const double* src = land + (size_t)b * BLOCK * p_dim + (size_t)row * p_dim;
            for (int c = 0; c < p_dim; ++c) Ps[(d_dim + c) * BLOCK] = src[c];
        } else {
            const double* src = la.theta + j * p_dim;
            for (int c = 0; c < p_dim; ++c) Ps[(d_dim + c) * BLOCK] = __ldg(src + c);
        }
        for (int c = 0; c < d_dim; ++c) Ps[c * BLOCK] = __ldg(la.d + i_d * d_dim + c);
        __syncthreads();                                   // landing buffer b^1 and `red` are free again
        {
            const long long nt = t + gridDim.x;
            if (nt < total && tid == 0) issue(nt, b ^ 1);
        }

        // ---- prologue ----
        run_tape<1>(vm, code_pro, img.n_pro);
        double xprev[NS];
#pragma unroll
        for (int n = 0; n < NS; ++n) xprev[n] = vm_ld(vm, tab_x0[n]);
        for (int q = 0; q < img.nq; ++q) S[(img.qe_slot + q) * BLOCK] = vm_ld(vm, tab_q0[q]);

        bool failed = false;
        unsigned my_iters = 0;

        for (int e = 0; e < img.nfe; ++e) {
            // controls of this finite element
            for (int c = 0; c < img.nce; ++c) {
                const int zi = tab_cmap[e * img.nce + c];
                double zv;
                if (la.z_mode == 0 || tab_zfix[zi]) zv = vm_ld(vm, tab_z[zi]);
                else if (la.z_mode == 1) zv = __ldg(la.z + zi);
                else if (la.z_mode == 2) zv = __ldg(la.z + i_d * img.nz + zi);
                else zv = __ldg(la.z + (i_d * la.n_t + j) * img.nz + zi);
                S[(img.ctrl_slot + c) * BLOCK] = zv;
            }
            run_tape<1>(vm, code_elem, img.n_elem);

            // Newton start: every collocation point at the element's initial state
#pragma unroll
            for (int n = 0; n < NS; ++n) {
                Xs[(3 * n + 0) * BLOCK] = xprev[n];
                Xs[(3 * n + 1) * BLOCK] = xprev[n];
                Xs[(3 * n + 2) * BLOCK] = xprev[n];
            }
            bool conv = false;
            for (int it = 0; it < img.max_it; ++it) {
                run_tape<3>(vm, code_pt, img.n_pt);

                double M[N][N], r[N], xv[3][NS];
#pragma unroll
                for (int kk = 0; kk < 3; ++kk)
#pragma unroll
                    for (int n = 0; n < NS; ++n) xv[kk][n] = Xs[(3 * n + kk) * BLOCK];
#pragma unroll
                for (int jj = 0; jj < 3; ++jj) {
#pragma unroll
                    for (int n = 0; n < NS; ++n) {
                        double lin = img.adot[jj + 1] * xprev[n];                    // adot[0][jj+1]
#pragma unroll
                        for (int kk = 0; kk < 3; ++kk) lin = fma(img.adot[4 * (kk + 1) + jj + 1], xv[kk][n], lin);
                        r[jj * NS + n] = fma(h, Fs[(3 * n + jj) * BLOCK], -lin);
#pragma unroll
                        for (int kk = 0; kk < 3; ++kk)
#pragma unroll
                            for (int m = 0; m < NS; ++m)
                                M[jj * NS + n][kk * NS + m] = (n == m) ? img.adot[4 * (kk + 1) + jj + 1] : 0.0;
                    }
                }
#pragma unroll
                for (int n = 0; n < NS; ++n)
#pragma unroll
                    for (int m = 0; m < NS; ++m)
                        if ((img.j_mask >> (n * NS + m)) & 1u) {
#pragma unroll
                            for (int jj = 0; jj < 3; ++jj)
                                M[jj * NS + n][jj * NS + m] =
                                    fma(-h, Js[(3 * (n * NS + m) + jj) * BLOCK], M[jj * NS + n][jj * NS + m]);
                        }

                lu_solve_inplace<N>(M, r);

                double dmax = 0.0, xmax = 1.0;
#pragma unroll
                for (int kk = 0; kk < 3; ++kk)
#pragma unroll
                    for (int n = 0; n < NS; ++n) {
                        const double dx = r[kk * NS + n];
                        const double xn = xv[kk][n] + dx;
                        Xs[(3 * n + kk) * BLOCK] = xn;
                        dmax = fmax(dmax, fabs(dx));
                        xmax = fmax(xmax, fabs(xn));
                        if (kk == 2) xv[2][n] = xn;
                    }
                if (!conv) ++my_iters;
                const bool bad = !(dmax <= DBL_MAX);          // inf or nan
                if (bad) failed = true;
                if (bad || dmax <= img.rtol * xmax) conv = true;
                cnt_warp_iters += 32;
                if (__all_sync(0xffffffffu, conv)) {
#pragma unroll
                    for (int n = 0; n < NS; ++n) xprev[n] = xv[2][n];
                    break;
                }
                if (it == img.max_it - 1) {
#pragma unroll
                    for (int n = 0; n < NS; ++n) xprev[n] = xv[2][n];
                }
            }
            if (!conv) failed = true;

            // quadrature states: x_q += h * sum_j bw_j fq(x_j) at the converged points
            if (img.nq > 0) {
                if (img.fq_state_dep) run_tape<3>(vm, code_pt, img.n_pt);
                for (int q = 0; q < img.nq; ++q) {
                    double acc = img.bw[0] * FQs[(3 * q + 0) * BLOCK];
                    acc = fma(img.bw[1], FQs[(3 * q + 1) * BLOCK], acc);
                    acc = fma(img.bw[2], FQs[(3 * q + 2) * BLOCK], acc);
                    S[(img.qe_slot + q) * BLOCK] = fma(h, acc, S[(img.qe_slot + q) * BLOCK]);
                }
            }
        }

        // ---- CQAs, indicator, outputs ----
#pragma unroll
        for (int n = 0; n < NS; ++n) S[(img.xe_slot + n) * BLOCK] = xprev[n];
        run_tape<1>(vm, code_g, img.n_g);
        bool feasible = !failed;
        for (int c = 0; c < img.ng; ++c) {
            double gv = vm_ld(vm, tab_g[c]);
            if (failed) gv = __longlong_as_double(0x7ff8000000000000LL);
            feasible = feasible && (gv <= 0.0);
            if (la.g_out && valid) la.g_out[(i_d * la.n_t + j) * img.ng + c] = gv;
        }
        if (la.ind_out && valid) la.ind_out[i_d * la.n_t + j] = feasible ? -0.0 : -1.0;

        double wv = 0.0;
        if (valid && feasible) wv = la.w ? __ldg(la.w + j) : 1.0 / (double)la.n_t;
        wv = warp_sum(wv);
        if (lane == 0) red[warp] = wv;
        if (valid) {
            cnt_iters += my_iters;
            cnt_failed += failed ? 1u : 0u;
            cnt_done += 1u;
        }
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < BLOCK / 32; ++i) s += red[i];
            la.partial[t] = s;
        }
    }

    // ---- counters: one atomic per warp ----
    if (la.counters) {
        const unsigned it_lo = warp_sum_u32((unsigned)cnt_iters);   // < 2^32 per CTA lifetime at these sizes
        const unsigned wi = warp_sum_u32((unsigned)(cnt_warp_iters / 32));
        const unsigned fl = warp_sum_u32(cnt_failed);
        const unsigned dn = warp_sum_u32(cnt_done);
        if (lane == 0) {
            atomicAdd(&la.counters[0], (unsigned long long)dn);
            atomicAdd(&la.counters[1], (unsigned long long)it_lo);
            atomicAdd(&la.counters[2], (unsigned long long)wi);
            atomicAdd(&la.counters[3], (unsigned long long)fl);
        }
    }
}

// Deterministic finish of the P(feasible) reduction: one thread per design point sums its tile
// partials in tile order.
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int n_tiles, long long n_d,
                                       double* __restrict__ P)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n_d) return;
    const double* p = partial + i * n_tiles;
    double s = 0.0;
    for (int t = 0; t < n_tiles; ++t) s += p[t];
    P[i] = s;
}

// FP64 FMA throughput probe: 8 independent chains per thread, no memory traffic.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace pydsb
