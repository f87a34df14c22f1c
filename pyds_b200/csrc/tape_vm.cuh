// Tape interpreter for the pyds_b200 feasibility kernels (sm_100a).
//
// One thread owns one (design point, theta) scenario.  Its register file lives in shared memory
// as S[slot][thread] (consecutive threads -> consecutive 8-byte words: every warp access is two
// conflict-free 128-byte wavefronts).  The instruction stream is block-uniform, so every fetch
// is a shared-memory broadcast and every branch of the dispatch is warp-uniform.
//
// Instruction word (host side: pyds_b200/tape.py):
//     op[7:0] | dst[21:8] | a[35:22] | b[49:36] | c[63:50]
// Operand reference (14 bit): bit 13 constant-pool index, bit 12 wide slot (three consecutive
// slots, one per collocation point), bits 11:0 index.  0x3FFF = structural zero.
//
// What it replaces in the reference: Pyomo's expression evaluation inside the GAMS/CONOPT solve
// that pyds/solver.py:50-55 launches for every design point (the model equations of
// run_twostage.py:49-65 are re-evaluated there on every NLP iteration).
#pragma once
#include <stdint.h>

namespace pydsb {

constexpr int BLOCK = 128;          // threads (= scenarios) per CTA tile

enum Op : unsigned {
    OP_MOV = 0, OP_NEG, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_FMA, OP_FMS, OP_FNMA, OP_SQR, OP_RCP, OP_EXP,
    OP_LOG, OP_SQRT, OP_SIN, OP_COS, OP_TANH, OP_POW, OP_ABS, OP_MIN, OP_MAX, OP_BC3, OP_COUNT
};

constexpr unsigned REF_CONST = 1u << 13;
constexpr unsigned REF_WIDE = 1u << 12;
constexpr unsigned REF_MASK = 0xFFFu;
constexpr unsigned REF_NONE = 0x3FFFu;

struct VM {
    double* S;              // slot file, already offset by the thread index
    const double* cpool;    // constant pool (shared memory, block-uniform); cpool[-1] == 0.0
};

struct Opnd {
    const double* p;
    int stride;             // in doubles, per collocation point
};

__device__ __forceinline__ Opnd vm_resolve(const VM& vm, unsigned ref)
{
    Opnd o;
    if (ref & REF_CONST) {
        o.p = (ref == REF_NONE) ? (vm.cpool - 1) : (vm.cpool + (ref & REF_MASK));
        o.stride = 0;
    } else {
        o.p = vm.S + (ref & REF_MASK) * BLOCK;
        o.stride = (ref & REF_WIDE) ? BLOCK : 0;
    }
    return o;
}

__device__ __forceinline__ double vm_ld(const VM& vm, unsigned ref)
{
    Opnd o = vm_resolve(vm, ref);
    return *o.p;
}

// Pivot reciprocal for the Newton LU: MUFU.RCP64H seed (20 bits) + two Newton steps with the
// second one folded (y*(1+e+e^2) form): relative error ~2^-60 before rounding.
__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double e2 = fma(e, e, e);
    return fma(y, e2, y);
}

#define PYDSB_UNARY(expr)                                                        \
    _Pragma("unroll") for (int w = 0; w < W; ++w) {                              \
        const double a = A.p[w * A.stride];                                      \
        D[w * dstride] = (expr);                                                 \
    }                                                                            \
    break;
#define PYDSB_BINARY(expr)                                                       \
    {                                                                            \
        const Opnd B = vm_resolve(vm, rb);                                       \
        _Pragma("unroll") for (int w = 0; w < W; ++w) {                          \
            const double a = A.p[w * A.stride], b = B.p[w * B.stride];           \
            D[w * dstride] = (expr);                                             \
        }                                                                        \
    }                                                                            \
    break;
#define PYDSB_TERNARY(expr)                                                      \
    {                                                                            \
        const Opnd B = vm_resolve(vm, rb);                                       \
        const Opnd C = vm_resolve(vm, rc);                                       \
        _Pragma("unroll") for (int w = 0; w < W; ++w) {                          \
            const double a = A.p[w * A.stride], b = B.p[w * B.stride], c = C.p[w * C.stride]; \
            D[w * dstride] = (expr);                                             \
        }                                                                        \
    }                                                                            \
    break;

// Run `n` instructions, each on W collocation points at once (W = 1 for the narrow tapes).
template <int W>
__device__ __forceinline__ void run_tape(const VM& vm, const uint64_t* __restrict__ code, int n)
{
    for (int pc = 0; pc < n; ++pc) {
        const uint64_t ins = code[pc];
        const unsigned lo = (unsigned)ins, hi = (unsigned)(ins >> 32);
        const unsigned op = lo & 0xFFu;
        const unsigned rd = (lo >> 8) & 0x3FFFu;
        const unsigned ra = (unsigned)(ins >> 22) & 0x3FFFu;
        const unsigned rb = (hi >> 4) & 0x3FFFu;
        const unsigned rc = (hi >> 18) & 0x3FFFu;
        double* D = vm.S + (rd & REF_MASK) * BLOCK;
        const int dstride = (rd & REF_WIDE) ? BLOCK : 0;
        const Opnd A = vm_resolve(vm, ra);
        switch (op) {
        case OP_MOV: PYDSB_UNARY(a)
        case OP_NEG: PYDSB_UNARY(-a)
        case OP_ADD: PYDSB_BINARY(a + b)
        case OP_SUB: PYDSB_BINARY(a - b)
        case OP_MUL: PYDSB_BINARY(a * b)
        case OP_DIV: PYDSB_BINARY(a / b)
        case OP_FMA: PYDSB_TERNARY(fma(a, b, c))
        case OP_FMS: PYDSB_TERNARY(fma(a, b, -c))
        case OP_FNMA: PYDSB_TERNARY(fma(-a, b, c))
        case OP_SQR: PYDSB_UNARY(a * a)
        case OP_RCP: PYDSB_UNARY(1.0 / a)
        case OP_EXP: PYDSB_UNARY(exp(a))
        case OP_LOG: PYDSB_UNARY(log(a))
        case OP_SQRT: PYDSB_UNARY(sqrt(a))
        case OP_SIN: PYDSB_UNARY(sin(a))
        case OP_COS: PYDSB_UNARY(cos(a))
        case OP_TANH: PYDSB_UNARY(tanh(a))
        case OP_POW: PYDSB_BINARY(pow(a, b))
        case OP_ABS: PYDSB_UNARY(fabs(a))
        case OP_MIN: PYDSB_BINARY(fmin(a, b))
        case OP_MAX: PYDSB_BINARY(fmax(a, b))
        case OP_BC3: {
            const double a = *A.p;
            D[0] = a; D[BLOCK] = a; D[2 * BLOCK] = a;
            break;
        }
        default: break;
        }
    }
}

#undef PYDSB_UNARY
#undef PYDSB_BINARY
#undef PYDSB_TERNARY

}  // namespace pydsb
