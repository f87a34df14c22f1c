// Shared device-side definitions of the pyds_b200 kernels: the executable image of a lowered model,
// launch arguments, mbarrier / bulk-TMA helpers, the register-resident LU, warp reductions.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "tape_vm.cuh"

namespace pydsb {

constexpr int MAX_BLOCKS = 4;          // blocks of the block-triangular state solve
constexpr int MAX_STATES = 8;          // dynamic states in total (each block holds at most 3)

constexpr int MAX_SENS_ENTRIES = 48;
struct SensEntry {
    unsigned short row, col;
    unsigned o[3];
};

struct ExecImage {
    // global-memory copies made by pydsb_model_create
    // (field order matters more than it should: shifting the offsets of the fields below by 8 or 20 bytes made ptxas
    //  spend 8 more registers on feas_kernel<1,2> and cost 5 % -- sens-only fields go to the end of the struct)
    const void* reserved_;
    const double* consts;      // constant pool
    const uint16_t* tabs;      // par[n_param] x0[ns] q0[nq] ztab[nz] zfix[nz] gtab[ng] cmap[nfe*nce] (+ sens tables)
    int n_pro, n_elem, n_pt, n_g, n_const, n_tabs, n_prog;
    int ns, nq, ng, d_dim, p_dim, nz, nfe, nce, n_units;
    int ctrl_slot, xe_slot, qe_slot, x_slot;
    int fq_state_dep;
    int max_it;
    unsigned zero_off;         // byte offset of a slot INIT keeps at 0.0 (0: the model needs none)
    int zero_slot;
    int spt;                   // scenarios per thread the offsets of this image are laid out for
    double rtol2;              // (relative Newton tolerance)^2
    double adot[16];
    double bw[3];
    // block-triangular state solve: block b owns blk_ns[b] states and instructions
    // [blk_pc0[b], blk_pc0[b] + blk_npt[b]) of the pre-decoded point tape; then the quadrature tape
    int n_blocks;
    int blk_ns[MAX_BLOCKS], blk_lin[MAX_BLOCKS], blk_pc0[MAX_BLOCKS], blk_npt[MAX_BLOCKS];
    unsigned blk_flops[MAX_BLOCKS];            // algorithmic flops of one Newton iteration of the block
    int q_pc0, q_npt;
    // byte offsets (per collocation point; equal offsets = narrow value) so that every access of the
    // compiled Newton code is LDS/STS [thread base + uniform offset]
    unsigned blk_x_o[MAX_BLOCKS][3][3];        // the block's states in the X region
    unsigned blk_f_o[MAX_BLOCKS][3][3];        // f of the block's states
    unsigned blk_j_o[MAX_BLOCKS][9][3];        // d f_r / d x_c within the block
    unsigned blk_jmask[MAX_BLOCKS];            // bit r*nsb+c: entry not structurally zero
    unsigned fq_o[8][3];
    unsigned fq_mask;
    // program 1 only (one block): the non-zero entries of df/dz (row = state, col = control function), then dfq/dx
    // (row = quadrature, col = state), then dfq/dz (row = quadrature, col = control function), with their byte
    // offsets at the three points; dg/dxe (ng x ns) and dg/dqe (ng x nq) are narrow references in `tabs` after cmap
    int n_fz, n_fqx, n_fqz;
    int tri;                   // df/dx is lower triangular in the state order (forward substitution over 3x3 systems)
    SensEntry sens_e[MAX_SENS_ENTRIES];
    double bigm_inv[8];        // 1 / M_k
    // shared-memory carve-up (bytes from the start of dynamic smem)
    int off_const, off_tabs, off_land, off_slots, smem_bytes;
    int seg[5];                // program 1: uint4 index where the pro / elem / point / g ops start, and the end
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
    double gtol[8];            // scenario feasible <=> g_k <= gtol[k] for every k (0: the plain rule g <= 0)
};

// Program 0 (feasibility): the whole per-scenario program -- parameter tape, element loop, Newton
// blocks, quadratures, CQA tape -- assembled by pydsb_model_create (assembler.h) into one pre-decoded
// instruction stream in the constant bank: 32 bytes per instruction,
//   word 0: { op | flags<<8 | aux<<16, dst, a, b }   word 1: { c, x1, x2, x3 }
// operands are byte offsets from the thread's base in the shared-memory register file, so that the
// fetch is a uniform LDCU and every access is LDS/STS [thread base + uniform offset (+ point * 1 KiB)].
constexpr int MAX_PROG = 768;
__constant__ uint4 c_prog[2 * MAX_PROG];
// Program 1 (feasibility + control sensitivities, sens_kernel): the data ops of its four tapes (pro | elem | point
// tape | g) in the same instruction format; segment boundaries in ExecImage::seg.  The Newton loop, the sensitivity
// systems and the objective statistics around them are hand-written (recourse_kernel.cuh).
constexpr int MAX_PROG_S = 512;
__constant__ uint4 c_prog_s[2 * MAX_PROG_S];

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
// STEP: the factors serve ONE Newton step (inexact Newton: the residual is exact, so the fixed point and the stop test do
// not change) -- the pivot reciprocals then carry one correction of the 20-bit seed (relative error 2^-40) instead of two.
template <int N, bool STEP = false>
__device__ __forceinline__ void lu_factor(double (&M)[N][N], double (&inv)[N])
{
#pragma unroll
    for (int k = 0; k < N; ++k) {
        inv[k] = STEP ? step_rcp(M[k][k]) : fast_rcp(M[k][k]);
#pragma unroll
        for (int i = k + 1; i < N; ++i) {
            const double l = M[i][k] * inv[k];
            M[i][k] = l;                                   // keep the multiplier: more right-hand sides may follow
#pragma unroll
            for (int c = k + 1; c < N; ++c) M[i][c] = fma(-l, M[k][c], M[i][c]);
        }
    }
}

template <int N>
__device__ __forceinline__ void lu_apply(const double (&M)[N][N], const double (&inv)[N], double (&r)[N])
{
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
        for (int i = k + 1; i < N; ++i) r[i] = fma(-M[i][k], r[k], r[i]);
#pragma unroll
    for (int k = N - 1; k >= 0; --k) {
        double s = r[k];
#pragma unroll
        for (int c = k + 1; c < N; ++c) s = fma(-M[k][c], r[c], s);
        r[k] = s * inv[k];
    }
}

template <int N, bool STEP = false>
__device__ __forceinline__ void lu_solve_inplace(double (&M)[N][N], double (&r)[N])
{
    double inv[N];
    lu_factor<N, STEP>(M, inv);
    lu_apply<N>(M, inv, r);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
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


// Everything transcendental goes through one out-of-line function (cold in the shipped models:
// the Arrhenius factors are parameter-only values, evaluated once per scenario).
__device__ __noinline__ double cold_op(unsigned op, double a, double b)
{
    switch (op) {
    case OP_EXP: return exp(a);
    case OP_LOG: return log(a);
    case OP_SQRT: return sqrt(a);
    case OP_SIN: return sin(a);
    case OP_COS: return cos(a);
    case OP_TANH: return tanh(a);
    case OP_POW: return pow(a, b);
    default: return a;
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
