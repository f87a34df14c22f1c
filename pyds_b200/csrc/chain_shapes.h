// Registered chain shapes: feas_kernel instantiations in which the structure of the fused element loop (poly_chain.cuh,
// ChainShape) is folded at compile time.  A model whose chain has another shape runs the generic instantiation (same
// results, the shape decoded per element).  To register one: `capi.assemble(blob)[1]`
// holds the two shape words of a lowered model; add a line here, rebuild.
//     X(SB, SQ, blocks held in registers)
#pragma once

// the two shape words of the A -> B -> C kinetics chain (tests/test_assembler_cpu.py pins them on the shipped models)
#define PYDSB_SHAPE_ABC_SB 0x18B980062ull
#define PYDSB_SHAPE_ABC_SQ 0x5C1Au

#define PYDSB_CHAIN_SHAPES(X)                                                                                           \
    /* A -> B -> C semi-batch kinetics, fixed feed (models.kinetics("design4" / "design6"); BASELINE configs C3, C5): */ \
    /* cA Newton {c0 narrow, c2 narrow}; cB linear with constant decay {c0 = narrow * cA^2, c1 narrow}: one inverse;  */ \
    /* quadratures u (element-constant integrand) and cC (narrow * cB)                                                */ \
    X(PYDSB_SHAPE_ABC_SB, PYDSB_SHAPE_ABC_SQ, 2)                                                                        \
    /* the same with the feed F_in a per-element control (models.kinetics("twostage"), run_twostage.py:19,79)         */ \
    X(PYDSB_SHAPE_ABC_SB | 4ull, PYDSB_SHAPE_ABC_SQ, 2)
