// Tape encoding shared by the host-side lowering (pyds_b200/tape.py), the assembler (assembler.h) and the cold
// paths of the kernels (reference tables are 14-bit references decoded with vm_ld).
//
// One thread owns one (design point, theta) scenario.  Its register file lives in shared memory
// as S[slot][thread] (consecutive threads -> consecutive 8-byte words: every warp access is two
// conflict-free 128-byte wavefronts).  The kernels do not interpret this encoding directly: pydsb_model_create
// assembles it into the pre-decoded programs of xvm_ops.cuh.
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

// Pivot reciprocal of a Newton STEP (lu_factor<N, true>): the bare MUFU seed, relative error 2^-20.  The residual of the
// step is exact, so the fixed point does not move; the step is off by ~2^-20 x conditioning of itself, which the caller's
// stop test accounts for (poly_chain.cuh).  PYDSB_STEP_RCP = 1: seed + one correction (2^-40), 0: fast_rcp (A/B builds).
#ifndef PYDSB_STEP_RCP
#define PYDSB_STEP_RCP 2
#endif
__device__ __forceinline__ double step_rcp(double x)
{
#if PYDSB_STEP_RCP == 0
    return fast_rcp(x);
#else
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if PYDSB_STEP_RCP == 1
    return fma(y, fma(-x, y, 1.0), y);
#else
    return y;
#endif
#endif
}

}  // namespace pydsb
