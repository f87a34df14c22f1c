// Three-wide point-tape interpreter of the sensitivity program (sens_kernel): operands addressed as
// [thread base + uniform byte offset], pre-decoded tape in the constant bank (c_pt_s).  The
// feasibility kernel runs the unified program of feas_kernel.cuh instead.
#pragma once
#include "common.cuh"

namespace pydsb {

// ---------------------------------------------------------------------------------------------
// Hot interpreter: the point tape, three collocation points per instruction, operands addressed
// as [thread base + uniform byte offset] (LDS/STS with a uniform-register offset).  Sixteen dense
// execution opcodes (XOP_*) so that the dispatch is one jump table; everything transcendental
// goes through one out-of-line cold function.
// ---------------------------------------------------------------------------------------------
enum XOp : unsigned {
    XOP_FMA = 0, XOP_FMS, XOP_FNMA, XOP_MUL, XOP_ADD, XOP_SUB, XOP_SQR, XOP_MOV, XOP_NEG, XOP_DIV, XOP_RCP,
    XOP_ABS, XOP_MIN, XOP_MAX, XOP_NOP, XOP_COLD
};

#define PT_LD(off) (*reinterpret_cast<const double*>(base + (off)))
#define PT_ST(off, v) (*reinterpret_cast<double*>(base + (off)) = (v))
// a narrow operand (same slot for the three points) is loaded once
#define PT_LOAD3(v, o0, o1, o2, narrow)                               \
    const double v##0 = PT_LD(o0);                                    \
    double v##1 = v##0, v##2 = v##0;                                  \
    if (!(narrow)) { v##1 = PT_LD(o1); v##2 = PT_LD(o2); }
#define PT_UN(expr)                                                   \
    {                                                                 \
        PT_LOAD3(a, q1.x, q1.y, q1.z, q0.x & 16u)                     \
        double a;                                                     \
        a = a0; const double r0 = (expr);                             \
        a = a1; const double r1 = (expr);                             \
        a = a2; const double r2 = (expr);                             \
        PT_ST(q0.y, r0); PT_ST(q0.z, r1); PT_ST(q0.w, r2);            \
    }                                                                 \
    break;
#define PT_BIN(expr)                                                  \
    {                                                                 \
        const uint4 q2 = c_pt[4 * pc + 2];                            \
        PT_LOAD3(a, q1.x, q1.y, q1.z, q0.x & 16u)                     \
        PT_LOAD3(b, q1.w, q2.x, q2.y, q0.x & 32u)                     \
        double a, b;                                                  \
        a = a0; b = b0; const double r0 = (expr);                     \
        a = a1; b = b1; const double r1 = (expr);                     \
        a = a2; b = b2; const double r2 = (expr);                     \
        PT_ST(q0.y, r0); PT_ST(q0.z, r1); PT_ST(q0.w, r2);            \
    }                                                                 \
    break;
#define PT_TER(expr)                                                  \
    {                                                                 \
        const uint4 q2 = c_pt[4 * pc + 2];                            \
        const uint4 q3 = c_pt[4 * pc + 3];                            \
        PT_LOAD3(a, q1.x, q1.y, q1.z, q0.x & 16u)                     \
        PT_LOAD3(b, q1.w, q2.x, q2.y, q0.x & 32u)                     \
        PT_LOAD3(c, q2.z, q2.w, q3.x, q0.x & 64u)                     \
        double a, b, c;                                               \
        a = a0; b = b0; c = c0; const double r0 = (expr);             \
        a = a1; b = b1; c = c1; const double r1 = (expr);             \
        a = a2; b = b2; c = c2; const double r2 = (expr);             \
        PT_ST(q0.y, r0); PT_ST(q0.z, r1); PT_ST(q0.w, r2);            \
    }                                                                 \
    break;

template <bool SENS>
__device__ __forceinline__ void run_point_tape(char* __restrict__ base, int pc0, int n)
{
    const uint4* const c_pt = c_pt_s;
    for (int pc = pc0; pc < pc0 + n; ++pc) {
        const uint4 q0 = c_pt[4 * pc];
        const uint4 q1 = c_pt[4 * pc + 1];
        switch (q0.x & 15u) {
        case XOP_FMA: PT_TER(fma(a, b, c))
        case XOP_FMS: PT_TER(fma(a, b, -c))
        case XOP_FNMA: PT_TER(fma(-a, b, c))
        case XOP_MUL: PT_BIN(a * b)
        case XOP_ADD: PT_BIN(a + b)
        case XOP_SUB: PT_BIN(a - b)
        case XOP_SQR: PT_UN(a * a)
        case XOP_MOV: PT_UN(a)
        case XOP_NEG: PT_UN(-a)
        case XOP_DIV: PT_BIN(a / b)
        case XOP_RCP: PT_UN(1.0 / a)
        case XOP_ABS: PT_UN(fabs(a))
        case XOP_MIN: PT_BIN(fmin(a, b))
        case XOP_MAX: PT_BIN(fmax(a, b))
        case XOP_NOP: break;
        case XOP_COLD: {
            const unsigned op = q0.x >> 8;
            PT_BIN(cold_op(op, a, b))
        }
        }
    }
}
#undef PT_LOAD3
#undef PT_UN
#undef PT_BIN
#undef PT_TER

// The narrow tapes (pro / elem / g) run once per scenario or element: keep one out-of-line copy.
__device__ __noinline__ void run_narrow_tape(double* S, const double* cpool, const uint64_t* code, int n)
{
    VM vm;
    vm.S = S;
    vm.cpool = cpool;
    run_tape<1>(vm, code, n);
}

#undef PT_LD
#undef PT_ST

}  // namespace pydsb
