// The instruction set of the per-scenario programs (feasibility: feas_kernel.cuh, sensitivities:
// recourse_kernel.cuh): opcodes, instruction layout, and the bodies of the data ops as macros.  A body expects, in
// scope:  base (char*, the thread's slot base), cz (const char*, constant pool - 8), c_prog (the program in the
// constant bank), pcw (uint4 index of the NEXT instruction), w0 (first uint4 of the current one), VecT / SPT / USP.
#pragma once
#include "common.cuh"

namespace pydsb {

// Opcodes, numbered by dispatch tier: the dispatch is a compare tree over uniform registers, so the ops
// that run once per Newton iteration sit at depth <= 3, the other common point ops at depth <= 5.
enum XCode : unsigned {
    // tier 1 (op < 4)
    X_POLY1 = 0, X_MUL_NW, X_NS1, X_ELEM_END,
    // tier 2 (op < 16)
    X_FMA_NWN, X_FMA_NWW, X_SQR_W, X_FMS_NWW, X_MUL_WW, X_MULSQ_NW, X_MUL1, X_FMA1, X_ADD1, X_NS1L, X_NS2, X_NB1,
    // tier 3: remaining three-wide point ops, by operand pattern (N narrow, W wide) ...
    X_ADD_NW, X_FMA_WWN, X_FMA_WWW, X_FMS_NWN, X_FMS_WWN, X_FMS_WWW, X_FNMA_NWN, X_FNMA_NWW, X_FNMA_WWN, X_FNMA_WWW, X_ADD_WW,
    X_SUB_WW, X_SUB_NW, X_SUB_WN, X_DIV_NW, X_DIV_WN, X_DIV_WW,
    X_NEG_W, X_RCP_W, X_ABS_W, X_MOV_W, X_BC_N, X_MIN_NW, X_MIN_WW, X_MAX_NW, X_MAX_WW, X_COLD_W,
    // ... scalar ops (pro / elem / g tapes; flags: operand a/b/c lives in the constant pool) ...
    X_FMS1, X_FNMA1, X_SUB1, X_DIV1, X_SQR1, X_NEG1, X_RCP1, X_ABS1, X_MOV1, X_MIN1, X_MAX1, X_COLD1,
    // ... structure
    X_INIT, X_LDCTRL, X_NB2, X_NB3, X_NS3, X_NS2L, X_NS3L, X_XEND, X_END,
    // ... the whole element loop of a chain of polynomial one-state blocks (poly_chain.cuh)
    X_PCHAIN,
    // ... chain programs only: r = fma(+-a, b, +-c), every FMA-class scalar op (flags bit 3: negate a, bit 4: negate c)
    X_FMAG,
    X_COUNT
};
constexpr unsigned X_TIER1 = 4, X_TIER2 = 16;

// Instruction layout (uint4 units; `pcw` indexes c_prog):
//   data ops    w0 = { op, dst, a, b }            w1 = { c, flags, cold opcode, - }                 length 2
//   NBk         w0 = { op, x_o[0..2] }                                                             length 2
//   NSk / NSkL  w0 = { op, target pcw, flops, jmask | max_it << 16 }
//               then x_o[k], f_o[k][3], j_o[k*k][3] packed word by word                          length 1 + ceil(k(4+3k)/4)
//   POLY1       w0 = { op, x_o, flops, max_it }   w1 = { c0, c1, c2, wide bits 0-2 | linear << 3 }
//   LDCTRL      w0 = { op, dst, control }
//   ELEM_END    w0 = { op, target pcw, n }        then n x { dst, o0, o1, o2 | scale, has_scale, -, - }: quadratures
//               to advance, optionally times a narrow scale factor                                 length 1 + 2n
//   PCHAIN      w0 = { op, nb | nq << 8, max_it }  then 3 words per block, 1 per quadrature (poly_chain.cuh)   length 2 + 3 nb + nq
__host__ __device__ constexpr int ns_words(int nsb) { return nsb + 3 * nsb + 3 * nsb * nsb; }
__host__ __device__ constexpr int ns_length(int nsb) { return 1 + (ns_words(nsb) + 3) / 4; }

constexpr unsigned US = BLOCK * 8u;                       // bytes between consecutive slots, per scenario of a thread

// The SPT scenarios of one thread, adjacent in the shared-memory register file: S[slot][thread][scenario].
template <int SPT>
struct alignas(8 * SPT) Vec {
    double v[SPT];
};

#define LDV(o) (*reinterpret_cast<const VecT*>(base + (o)))
#define LDVP(o, p) (*reinterpret_cast<const VecT*>(base + (o) + (p) * USP))
#define STV(o, val) (*reinterpret_cast<VecT*>(base + (o)) = (val))
#define STVP(o, p, val) (*reinterpret_cast<VecT*>(base + (o) + (p) * USP) = (val))


// operand loaders of the point ops: N = narrow (one load), W = wide (three loads, one slot apart)
#define XLD_N(v, o) const VecT v##0 = LDV(o); const VecT v##1 = v##0, v##2 = v##0;
#define XLD_W(v, o) const VecT v##0 = LDVP(o, 0), v##1 = LDVP(o, 1), v##2 = LDVP(o, 2);
#define XST3(o) STVP(o, 0, r0); STVP(o, 1, r1); STVP(o, 2, r2);
#define X_TER(PA, PB, PC, expr)                                       \
    {                                                                 \
        const unsigned oc = c_prog[pcw - 1].x;                        \
        XLD_##PA(a, w0.z) XLD_##PB(b, w0.w) XLD_##PC(c, oc)           \
        VecT r0, r1, r2;                                              \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s) {             \
            double a, b, c;                                           \
            a = a0.v[s]; b = b0.v[s]; c = c0.v[s]; r0.v[s] = (expr);  \
            a = a1.v[s]; b = b1.v[s]; c = c1.v[s]; r1.v[s] = (expr);  \
            a = a2.v[s]; b = b2.v[s]; c = c2.v[s]; r2.v[s] = (expr);  \
        }                                                             \
        XST3(w0.y)                                                    \
    }                                                                 \
    break;
#define X_BIN(PA, PB, expr)                                           \
    {                                                                 \
        XLD_##PA(a, w0.z) XLD_##PB(b, w0.w)                           \
        VecT r0, r1, r2;                                              \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s) {             \
            double a, b;                                              \
            a = a0.v[s]; b = b0.v[s]; r0.v[s] = (expr);               \
            a = a1.v[s]; b = b1.v[s]; r1.v[s] = (expr);               \
            a = a2.v[s]; b = b2.v[s]; r2.v[s] = (expr);               \
        }                                                             \
        XST3(w0.y)                                                    \
    }                                                                 \
    break;
#define X_UN(PA, expr)                                                \
    {                                                                 \
        XLD_##PA(a, w0.z)                                             \
        VecT r0, r1, r2;                                              \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s) {             \
            double a;                                                 \
            a = a0.v[s]; r0.v[s] = (expr);                            \
            a = a1.v[s]; r1.v[s] = (expr);                            \
            a = a2.v[s]; r2.v[s] = (expr);                            \
        }                                                             \
        XST3(w0.y)                                                    \
    }                                                                 \
    break;
// scalar ops: an operand is a slot of this thread or an entry of the (block-uniform) constant pool
#define X1_OPND(n, bit, off)                                                             \
    VecT n##v;                                                                           \
    if (fl & (bit)) {                                                                    \
        const double kc = *reinterpret_cast<const double*>(cz + (off));                  \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s) n##v.v[s] = kc;                  \
    } else n##v = LDV(off);
#define X1_FL const uint4 w1 = c_prog[pcw - 1]; const unsigned fl = w1.y;
#define X1_TER(expr)                                                                     \
    {                                                                                    \
        X1_FL X1_OPND(a, 1u, w0.z) X1_OPND(b, 2u, w0.w) X1_OPND(c, 4u, w1.x)             \
        VecT rv;                                                                         \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s) {                                \
            const double a = av.v[s], b = bv.v[s], c = cv.v[s];                          \
            rv.v[s] = (expr);                                                            \
        }                                                                                \
        STV(w0.y, rv);                                                                   \
    }                                                                                    \
    break;
#define X1_BIN(expr)                                                                     \
    {                                                                                    \
        X1_FL X1_OPND(a, 1u, w0.z) X1_OPND(b, 2u, w0.w)                                  \
        VecT rv;                                                                         \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s) {                                \
            const double a = av.v[s], b = bv.v[s];                                       \
            rv.v[s] = (expr);                                                            \
        }                                                                                \
        STV(w0.y, rv);                                                                   \
    }                                                                                    \
    break;
#define X1_UN(expr)                                                                      \
    {                                                                                    \
        X1_FL X1_OPND(a, 1u, w0.z)                                                       \
        VecT rv;                                                                         \
        _Pragma("unroll") for (int s = 0; s < SPT; ++s) {                                \
            const double a = av.v[s];                                                    \
            rv.v[s] = (expr);                                                            \
        }                                                                                \
        STV(w0.y, rv);                                                                   \
    }                                                                                    \
    break;

// transcendental point op, 3-wide; flag bit 0/1 (w1.y): operand a/b is wide; w1.z = the tape opcode
#define X_COLD_W_BODY                                                                    \
    {                                                                                    \
        const uint4 w1 = c_prog[pcw - 1];                                                \
        const unsigned sa = (w1.y & 1u) ? USP : 0u, sb = (w1.y & 2u) ? USP : 0u;         \
        _Pragma("unroll 1") for (unsigned p = 0; p < 3; ++p) {                           \
            const VecT a = LDV(w0.z + p * sa), b = LDV(w0.w + p * sb);                   \
            VecT r;                                                                      \
            _Pragma("unroll") for (int s = 0; s < SPT; ++s) r.v[s] = cold_op(w1.z, a.v[s], b.v[s]); \
            STV(w0.y + p * USP, r);                                                      \
        }                                                                                \
    }                                                                                    \
    break;
// every data op (point ops by operand pattern, scalar ops) as switch cases, split like the dispatch tiers: the
// common ones (op < X_TIER2) and the rest -- for interpreters without structure ops (sens_kernel)
#define X_DATA_OP_CASES_HOT                                               \
    case X_MUL_NW: X_BIN(N, W, a * b)                                     \
    case X_FMA_NWN: X_TER(N, W, N, fma(a, b, c))                          \
    case X_FMA_NWW: X_TER(N, W, W, fma(a, b, c))                          \
    case X_SQR_W: X_UN(W, a * a)                                          \
    case X_FMS_NWW: X_TER(N, W, W, fma(a, b, -c))                         \
    case X_MUL_WW: X_BIN(W, W, a * b)                                     \
    case X_MULSQ_NW: X_BIN(N, W, a * b * b)                               \
    case X_MUL1: X1_BIN(a * b)                                            \
    case X_FMA1: X1_TER(fma(a, b, c))                                     \
    case X_ADD1: X1_BIN(a + b)
#define X_DATA_OP_CASES_REST                                              \
    case X_ADD_NW: X_BIN(N, W, a + b)                                     \
    case X_FMA_WWN: X_TER(W, W, N, fma(a, b, c))                          \
    case X_FMA_WWW: X_TER(W, W, W, fma(a, b, c))                          \
    case X_FMS_NWN: X_TER(N, W, N, fma(a, b, -c))                         \
    case X_FMS_WWN: X_TER(W, W, N, fma(a, b, -c))                         \
    case X_FMS_WWW: X_TER(W, W, W, fma(a, b, -c))                         \
    case X_FNMA_NWN: X_TER(N, W, N, fma(-a, b, c))                        \
    case X_FNMA_NWW: X_TER(N, W, W, fma(-a, b, c))                        \
    case X_FNMA_WWN: X_TER(W, W, N, fma(-a, b, c))                        \
    case X_FNMA_WWW: X_TER(W, W, W, fma(-a, b, c))                        \
    case X_ADD_WW: X_BIN(W, W, a + b)                                     \
    case X_SUB_WW: X_BIN(W, W, a - b)                                     \
    case X_SUB_NW: X_BIN(N, W, a - b)                                     \
    case X_SUB_WN: X_BIN(W, N, a - b)                                     \
    case X_DIV_NW: X_BIN(N, W, a / b)                                     \
    case X_DIV_WN: X_BIN(W, N, a / b)                                     \
    case X_DIV_WW: X_BIN(W, W, a / b)                                     \
    case X_NEG_W: X_UN(W, -a)                                             \
    case X_RCP_W: X_UN(W, 1.0 / a)                                        \
    case X_ABS_W: X_UN(W, fabs(a))                                        \
    case X_MOV_W: X_UN(W, a)                                              \
    case X_BC_N: X_UN(N, a)                                               \
    case X_MIN_NW: X_BIN(N, W, fmin(a, b))                                \
    case X_MIN_WW: X_BIN(W, W, fmin(a, b))                                \
    case X_MAX_NW: X_BIN(N, W, fmax(a, b))                                \
    case X_MAX_WW: X_BIN(W, W, fmax(a, b))                                \
    case X_COLD_W: X_COLD_W_BODY                                          \
    case X_FMS1: X1_TER(fma(a, b, -c))                                    \
    case X_FNMA1: X1_TER(fma(-a, b, c))                                   \
    case X_SUB1: X1_BIN(a - b)                                            \
    case X_DIV1: X1_BIN(a / b)                                            \
    case X_SQR1: X1_UN(a * a)                                             \
    case X_NEG1: X1_UN(-a)                                                \
    case X_RCP1: X1_UN(1.0 / a)                                           \
    case X_ABS1: X1_UN(fabs(a))                                           \
    case X_MOV1: X1_UN(a)                                                 \
    case X_MIN1: X1_BIN(fmin(a, b))                                       \
    case X_MAX1: X1_BIN(fmax(a, b))                                       \
    case X_COLD1: X1_BIN(cold_op(w1.z, a, b))

}  // namespace pydsb
