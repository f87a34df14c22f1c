// Host side of the feasibility interpreter: assemble the lowered tapes of a model blob (pro / elem /
// point tapes per block / quadrature tape / g -- pyds_b200/tape.py) into the ONE pre-decoded program
// feas_kernel runs per scenario (layout and opcodes: feas_kernel.cuh).  Every operand becomes a byte
// offset from the thread's base in the shared-memory register file; every point op is specialised on
// its operand pattern (narrow / wide), commutative operands are ordered narrow first.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <string>
#include <vector>

#include "feas_kernel.cuh"

namespace pydsb {

// A one-state block lowered to the coefficients of  f = c0 + c1 x + c2 x^2  (byte offsets; narrow or wide)
struct PolyBlock {
    bool is_poly = false;
    bool present[3] = {false, false, false}, wide[3] = {false, false, false};
    uint32_t off[3] = {0, 0, 0};
};

struct AsmInput {
    const uint64_t *pro, *elem, *g, *pt;       // 64-bit tape words (tape_vm.cuh encoding)
    int n_pro, n_elem, n_g;
    int n_blocks;
    const int *blk_npt, *blk_ns, *blk_lin;
    const unsigned* blk_flops;
    int q_npt;
    int nce, ctrl_slot, nq, qe_slot, n_units, nfe;
    const uint16_t* cmap;                      // [nfe][nce] z index of every control on every finite element
    const uint16_t* fq_refs;                   // [nq] reference of every quadrature integrand
    int spt;                                   // scenarios per thread: slot stride = spt * US bytes
    const PolyBlock* poly;                     // [n_blocks]
    const ExecImage* img;                      // resolved per-block offsets: blk_x_o, blk_f_o, blk_j_o, blk_jmask
    bool allow_chain = false;                  // fuse the element loop into one X_PCHAIN when the model has that shape
    // reference tables (chain programs resolve them into the descriptor: no table walking in the kernel)
    const uint16_t *x0_refs = nullptr, *q0_refs = nullptr, *par_refs = nullptr, *g_refs = nullptr, *z_refs = nullptr;
    int n_par = 0, ng = 0, xe_slot = 0, nz = 0;
};

class Assembler {
public:
    std::vector<uint32_t> words;               // 8 per instruction
    std::string err;
    int chain_nb = 0;                          // blocks of the fused chain (0: the element loop is interpreted)
    ChainDesc chain{};                         // its descriptor (poly_chain.cuh): what the kernel reads from the constant bank
    int chain_units = 0;                       // register-file units a chain program touches after compaction (0: not compacted)
    bool compact = true;                       // renumber the units of a chain program by liveness (compact_units)

    int pc() const { return (int)(words.size() / 4); }      // uint4 index of the next instruction

    bool assemble(const AsmInput& in)
    {
        in_ = &in;
        // controls that map every finite element to the same z entry (fixed profiles, constant-in-time controls):
        // the control loads and the elem tape are element invariant and run once, before the loop
        bool invariant = true;
        for (int e = 1; e < in.nfe; ++e)
            for (int c = 0; c < in.nce; ++c) invariant = invariant && in.cmap[e * in.nce + c] == in.cmap[c];
        std::vector<uint32_t> chain;
        int nb = 0;
        chain_mode_ = in.allow_chain && try_chain(in, invariant, chain, nb);
        for (int i = 0; i < in.n_pro; ++i)
            if (!scalar_op(in.pro[i])) return false;
        if (!chain_mode_) emit(X_INIT);                    // a chain reads the initial values where they are
        auto ctrl_and_elem = [&]() -> bool {
            for (int c = 0; c < in.nce; ++c) {
                const uint32_t dst = (uint32_t)(in.ctrl_slot + c) * US * (uint32_t)in.spt;
                // chain programs: every scalar op loads its a / b operands before it looks at the opcode -- keep them valid
                if (chain_mode_) { emit(X_LDCTRL, dst, 0u, 0u, 0u, 0u, (uint32_t)c); nread_.push_back(0); }
                else emit(X_LDCTRL, dst, (uint32_t)c);
            }
            for (int i = 0; i < in.n_elem; ++i)
                if (!scalar_op(in.elem[i])) return false;
            return true;
        };
        if (chain_mode_) {
            // scalar ops | PCHAIN | XEND | scalar ops | END  (the shape feas_kernel<.., CHAIN> relies on)
            if (invariant && !ctrl_and_elem()) return false;
            std::vector<uint32_t> pre, post;
            pre.swap(words);
            for (int i = 0; i < in.n_g; ++i)
                if (!scalar_op(in.g[i])) return false;
            post.swap(words);
            if (compact && !compact_units(in, pre, post)) return false;
            words.swap(pre);
            this->chain.pre_end = (unsigned)pc();
            chain_words(chain);
            words.insert(words.end(), chain.begin(), chain.end());
            chain_nb = nb;
            emit(X_XEND);
            this->chain.g_begin = (unsigned)pc();
            words.insert(words.end(), post.begin(), post.end());
            this->chain.g_end = (unsigned)pc();
            emit(X_END);
            this->chain.pre_run0 = fmag_runs(0, this->chain.pre_end);
            this->chain.g_run0 = fmag_runs(this->chain.g_begin, this->chain.g_end);
            if (pc() > 2 * MAX_PROG - 2) return fail("program longer than the constant-bank window");
            return true;
        }
        if (invariant && !ctrl_and_elem()) return false;
        const int l_elem = pc();
        if (!invariant && !ctrl_and_elem()) return false;
        int pos = 0;
        for (int b = 0; b < in.n_blocks; ++b) {
            const int nsb = in.blk_ns[b];
            if (nsb < 1 || nsb > 3) return fail("block size out of range");
            static const uint32_t nb_code[3] = {X_NB1, X_NB2, X_NB3}, ns_code[3] = {X_NS1, X_NS2, X_NS3},
                                  nsl_code[3] = {X_NS1L, X_NS2L, X_NS3L};
            const ExecImage& g = *in.img;
            if (g.max_it < 1 || g.max_it > 0xFFFF) return fail("Newton iteration cap out of range");
            if (in.poly[b].is_poly) {
                // the coefficients that depend on earlier blocks' states (once per element), then the whole solve
                {
                    std::vector<unsigned> live;                      // slots POLY1 reads: the coefficients
                    for (int k = 0; k < 3; ++k)
                        if (in.poly[b].present[k]) live.push_back(in.poly[b].off[k] / (US * (uint32_t)in.spt));
                    if (!point_tape(in.pt + pos, in.blk_npt[b], live)) return false;
                    pos += in.blk_npt[b];
                }
                const PolyBlock& pb = in.poly[b];
                uint32_t off[3], flags = in.blk_lin[b] ? 8u : 0u;
                for (int k = 0; k < 3; ++k) {
                    if (!pb.present[k] && !g.zero_off) return fail("model needs the zero slot");
                    off[k] = pb.present[k] ? pb.off[k] : g.zero_off;
                    if (pb.present[k] && pb.wide[k]) flags |= 1u << k;
                }
                // c0 = narrow * wide^2 computed by the block's own coefficient tape (the MULSQ just emitted) and read by
                // nothing else: POLY1 forms it in registers (flag 16, a third word {narrow, wide}) -- one dispatch and a
                // store / reload of three values less per element (second-order rate k * c^2 feeding the next species)
                bool fused = false;
                const size_t nw = words.size();
#ifndef PYDSB_NO_C0SQ                                         /* development switch: A/B of the fusion */
                if (pb.present[0] && pb.wide[0] && in.blk_npt[b] >= 2 && nw >= 8 && words[nw - 8] == (uint32_t)X_MULSQ_NW &&
                    words[nw - 7] == pb.off[0] && !reads_units_elsewhere(in, pb.off[0] / (US * (uint32_t)in.spt), pos - 2, b)) {
                    const uint32_t a_off = words[nw - 6], b_off = words[nw - 5];
                    words.resize(nw - 8);
                    emit(X_POLY1, g.blk_x_o[b][0][0], in.blk_flops[b], (uint32_t)g.max_it, 0u, off[1], off[2], flags | 16u);
                    const uint32_t w2[4] = {a_off, b_off, 0u, 0u};
                    words.insert(words.end(), w2, w2 + 4);
                    fused = true;
                }
#endif
                if (!fused)
                    emit(X_POLY1, g.blk_x_o[b][0][0], in.blk_flops[b], (uint32_t)g.max_it, off[0], off[1], off[2], flags);
                continue;
            }
            emit(nb_code[nsb - 1], g.blk_x_o[b][0][0], nsb > 1 ? g.blk_x_o[b][1][0] : 0u, nsb > 2 ? g.blk_x_o[b][2][0] : 0u);
            const int l_blk = pc();
            {
                std::vector<unsigned> live;                          // slots NS reads: f and J of the block
                const uint32_t usp = US * (uint32_t)in.spt;
                for (int n = 0; n < nsb; ++n) live.push_back(g.blk_f_o[b][n][0] / usp);
                for (int i = 0; i < nsb * nsb; ++i)
                    if ((g.blk_jmask[b] >> i) & 1u) live.push_back(g.blk_j_o[b][i][0] / usp);
                if (!point_tape(in.pt + pos, in.blk_npt[b], live)) return false;
                pos += in.blk_npt[b];
            }
            // NS: first uint4, then the operand table x_o[k] | f_o[k][3] | j_o[k*k][3], padded to whole uint4s
            const uint32_t head[4] = {in.blk_lin[b] ? nsl_code[nsb - 1] : ns_code[nsb - 1], (uint32_t)l_blk, in.blk_flops[b],
                                      (g.blk_jmask[b] & 0xFFFFu) | ((uint32_t)g.max_it << 16)};
            words.insert(words.end(), head, head + 4);
            std::vector<uint32_t> tab;
            for (int n = 0; n < nsb; ++n) tab.push_back(g.blk_x_o[b][n][0]);
            for (int n = 0; n < nsb; ++n)
                for (int p3 = 0; p3 < 3; ++p3) tab.push_back(g.blk_f_o[b][n][p3]);
            for (int i = 0; i < nsb * nsb; ++i)
                for (int p3 = 0; p3 < 3; ++p3) {
                    // a 1-state block always loads its Jacobian entry: structurally zero -> the zero slot
                    const bool zero = nsb == 1 && !(g.blk_jmask[b] & 1u);
                    if (zero && !g.zero_off) return fail("model needs the zero slot");
                    tab.push_back(zero ? g.zero_off : g.blk_j_o[b][i][p3]);
                }
            tab.resize((size_t)(ns_length(nsb) - 1) * 4, 0u);
            words.insert(words.end(), tab.begin(), tab.end());
        }
        // quadrature tape and ELEM_END (advances the quadratures with the Radau weights, then loops).  An integrand that
        // is `narrow * wide` computed by the tape's own MUL and read by nothing else is folded into its ELEM_END entry
        // as a scale factor (mass-action rates: k * c): one dispatch and a store / reload less per element.
        {
            std::vector<uint64_t> qt(in.pt + pos, in.pt + pos + in.q_npt);
            pos += in.q_npt;
            std::vector<uint32_t> quad;                              // 8 words per entry
            for (int q = 0; q < in.nq; ++q) {
                const unsigned r = in.fq_refs[q];
                if (r == REF_NONE) continue;
                if (r & REF_CONST) return fail("constant reference in the quadrature table");
                if (!check_slot(r)) return false;
                uint32_t o = slot_off(r), st = (r & REF_WIDE) ? US * (uint32_t)in.spt : 0u, scale = 0, has_scale = 0;
                if (r & REF_WIDE) {
                    int producer = -1, readers = 0, others = 0;
                    for (size_t i = 0; i < qt.size(); ++i) {
                        unsigned op, refs[4];
                        split(qt[i], op, refs);
                        if ((refs[0] & REF_MASK) == (r & REF_MASK)) producer = (int)i;
                        for (int k = 1; k <= arity(op); ++k) readers += (refs[k] & REF_MASK) == (r & REF_MASK) && !(refs[k] & REF_CONST);
                    }
                    for (int q2 = 0; q2 < in.nq; ++q2) others += q2 != q && in.fq_refs[q2] == r;
                    if (producer >= 0 && readers == 0 && others == 0) {
                        unsigned op, refs[4];
                        split(qt[producer], op, refs);
                        unsigned na = refs[1], wb = refs[2];
                        if (op == OP_MUL && (na & REF_WIDE) && !(wb & REF_WIDE)) { const unsigned t = na; na = wb; wb = t; }
                        if (op == OP_MUL && !(na & (REF_WIDE | REF_CONST)) && (wb & REF_WIDE) && !(wb & REF_CONST) &&
                            check_slot(na) && check_slot(wb)) {
                            o = slot_off(wb); scale = slot_off(na); has_scale = 1;
                            qt.erase(qt.begin() + producer);
                        }
                    }
                }
                const uint32_t v[8] = {(uint32_t)(in.qe_slot + q) * US * (uint32_t)in.spt, o, o + st, o + 2 * st, scale, has_scale, 0, 0};
                quad.insert(quad.end(), v, v + 8);
            }
            if (!point_tape(qt.data(), (int)qt.size(), std::vector<unsigned>())) return false;
            const uint32_t head[4] = {X_ELEM_END, (uint32_t)l_elem, (uint32_t)(quad.size() / 8), 0u};
            words.insert(words.end(), head, head + 4);
            words.insert(words.end(), quad.begin(), quad.end());
        }
        emit(X_XEND);
        for (int i = 0; i < in.n_g; ++i)
            if (!scalar_op(in.g[i])) return false;
        emit(X_END);
        if (pc() > 2 * MAX_PROG - 2) return fail("program longer than the constant-bank window");
        return true;
    }

    // Program 1 (sens_kernel): the data ops of the four tapes back to back; seg[k] = where tape k starts, seg[4] = end
    bool assemble_segments(const AsmInput& in, int n_pt, int (&seg)[5])
    {
        in_ = &in;
        seg[0] = pc();
        for (int i = 0; i < in.n_pro; ++i)
            if (!scalar_op(in.pro[i])) return false;
        seg[1] = pc();
        for (int i = 0; i < in.n_elem; ++i)
            if (!scalar_op(in.elem[i])) return false;
        seg[2] = pc();
        for (int i = 0; i < n_pt; ++i)                           // (no peephole: sens_kernel reads most of these slots)
            if (!point_op(in.pt[i])) return false;
        seg[3] = pc();
        for (int i = 0; i < in.n_g; ++i)
            if (!scalar_op(in.g[i])) return false;
        seg[4] = pc();
        if (pc() > 2 * MAX_PROG_S) return fail("sensitivity program longer than the constant-bank window");
        return true;
    }

private:
    const AsmInput* in_ = nullptr;
    bool chain_mode_ = false;                  // chain programs: FMAG scalar ops, SPT-wide constant pool
    std::vector<unsigned char> nread_;         // chain programs: operand fields (a, b, c) every scalar op really reads

    // Scalar section [begin, end) of a chain program (every op two uint4 long): the FMAG ops that directly follow another
    // op are counted into that op's w1.w -- the kernel runs them in a loop without looking at opcodes.  Returns the
    // length of the run at the head of the section.
    unsigned fmag_runs(unsigned begin, unsigned end)
    {
        unsigned run0 = 0;
        long long last = -1;                                          // uint4 index of the last non-FMAG op
        for (unsigned p4 = begin; p4 < end; p4 += 2) {
            if (words[(size_t)p4 * 4] == (uint32_t)X_FMAG) {
                if (last < 0) ++run0;
                else ++words[(size_t)(last + 1) * 4 + 3];
            } else {
                last = p4;
                words[(size_t)(p4 + 1) * 4 + 3] = 0u;
            }
        }
        return run0;
    }

    // ---- fused polynomial chain (poly_chain.cuh) ----------------------------------------------------------------
    struct Term {
        bool present = false, has_n = false, has_ctrl = false, neg = false;
        uint32_t n_off = 0, ctrl = 0, p = 0, src = 0;
        uint32_t meta() const { return present ? (chain_meta(has_n, has_ctrl, ctrl, p, src) | (neg ? CT_NEG : 0u)) : 0u; }
    };
    struct ElemVal {                           // value of the elem tape:  (narrow slot a, or 1) * control c
        unsigned dst;
        bool has_a;
        unsigned a;
        unsigned c;
    };

    // Is `ref` the wide reference of a dynamic state's collocation points?  -> the block that owns the state
    static int state_block(const AsmInput& in, unsigned ref)
    {
        if ((ref & REF_CONST) || !(ref & REF_WIDE)) return -1;
        const ExecImage& g = *in.img;
        const uint32_t off = (ref & REF_MASK) * US * (uint32_t)in.spt;
        for (int b = 0; b < in.n_blocks; ++b)
            if (in.blk_ns[b] == 1 && g.blk_x_o[b][0][0] == off) return b;
        return -1;
    }

    // where an initial value lives: the constant pool of a chain program (SPT-wide entries after the three specials
    // 1.0, -0.0, 0.0) or a slot
    static bool source_ref(const AsmInput& in, unsigned ref, uint32_t& off, uint32_t& is_const)
    {
        const uint32_t cw = 8u * (uint32_t)in.spt;
        if (ref & REF_WIDE && !(ref & REF_CONST)) return false;
        if (ref & REF_CONST) {
            is_const = 1u;
            off = (ref == REF_NONE) ? 2u * cw : ((ref & REF_MASK) + 3u) * cw;
        } else {
            is_const = 0u;
            off = (ref & REF_MASK) * US * (uint32_t)in.spt;
        }
        return true;
    }

    // narrow factor: a plain persistent slot, or (controls that differ between the elements) a control value / an
    // elem-tape value `a * control`
    static bool narrow_factor(const AsmInput& in, unsigned ref, bool invariant, const std::vector<ElemVal>& ev, Term& t)
    {
        if (ref & (REF_CONST | REF_WIDE)) return false;
        const unsigned idx = ref & REF_MASK;
        if ((int)idx >= in.img->x_slot) return false;                  // not in the persistent region
        const uint32_t usp = US * (uint32_t)in.spt;
        if (!invariant) {
            if ((int)idx >= in.ctrl_slot && (int)idx < in.ctrl_slot + in.nce) {
                t.has_ctrl = true; t.ctrl = idx - (unsigned)in.ctrl_slot;
                return true;
            }
            for (const ElemVal& v : ev)
                if (v.dst == idx) {
                    t.has_ctrl = true; t.ctrl = v.c;
                    t.has_n = v.has_a; t.n_off = v.a * usp;
                    return true;
                }
        }
        t.has_n = true; t.n_off = idx * usp;
        return true;
    }

    // A wide value of the tape `code[0..n)`, as it stands before instruction `before`, as ONE term
    //     +-narrow * (state of an earlier block)^p:
    // the state itself (p = 1), SQR of it (p = 2), MUL narrow x such a value, NEG of such a value.  Slots are reused inside
    // a tape, so the producer of a reference is the LAST write before its reader.  Marks the tape ops the term stands for.
    static bool wide_value(const AsmInput& in, const uint64_t* code, int before, unsigned ref, bool invariant,
                           const std::vector<ElemVal>& ev, std::vector<char>& used, Term& t)
    {
        int at = -1;
        for (int i = before - 1; i >= 0 && at < 0; --i) {
            unsigned op, refs[4];
            split(code[i], op, refs);
            if (refs[0] == ref) at = i;
        }
        if (at < 0) {                                                  // not computed here: a state's collocation points
            const int b = state_block(in, ref);
            if (b < 0) return false;
            t.p = 1; t.src = (uint32_t)b;
            return true;
        }
        unsigned op, refs[4];
        split(code[at], op, refs);
        used[at] = 1;
        if (op == OP_NEG) {
            if (!wide_value(in, code, at, refs[1], invariant, ev, used, t)) return false;
            t.neg = !t.neg;
            return true;
        }
        if (op == OP_SQR) {
            if (!wide_value(in, code, at, refs[1], invariant, ev, used, t)) return false;
            if (t.p != 1 || t.has_n || t.has_ctrl || t.neg) return false;             // only the square of a bare state
            t.p = 2;
            return true;
        }
        if (op == OP_MUL) {
            unsigned na = refs[1], wb = refs[2];
            if ((na & REF_WIDE) && !(wb & REF_WIDE)) { const unsigned s = na; na = wb; wb = s; }
            if ((na & (REF_WIDE | REF_CONST)) || !(wb & REF_WIDE) || (wb & REF_CONST)) return false;
            if (!wide_value(in, code, at, wb, invariant, ev, used, t)) return false;
            if (t.has_n || t.has_ctrl) return false;                   // one narrow factor per term
            return narrow_factor(in, na, invariant, ev, t);
        }
        return false;
    }
    static bool wide_term(const AsmInput& in, const uint64_t* code, int n, unsigned ref, int own_block, bool invariant,
                          const std::vector<ElemVal>& ev, std::vector<char>& used, Term& t)
    {
        if (!wide_value(in, code, n, ref, invariant, ev, used, t)) return false;
        if ((int)t.src >= own_block) return false;                     // only states of earlier blocks
        t.present = true;
        return true;
    }

    // The element loop as one X_PCHAIN instruction, or false when the model does not have that shape.
    bool try_chain(const AsmInput& in, bool invariant, std::vector<uint32_t>& out, int& nb_out)
    {
        const ExecImage& g = *in.img;
        const uint32_t usp = US * (uint32_t)in.spt;
        if (in.n_blocks < 1 || in.n_blocks > CHAIN_MAXB) return false;
        for (int b = 0; b < in.n_blocks; ++b)
            if (!in.poly[b].is_poly || in.blk_ns[b] != 1) return false;
        // controls that differ between the elements: every elem-tape value must be  narrow * control
        std::vector<ElemVal> ev;
        if (!invariant) {
            for (int i = 0; i < in.n_elem; ++i) {
                unsigned op, refs[4];
                split(in.elem[i], op, refs);
                if (op != OP_MUL) return false;
                ElemVal v{refs[0] & REF_MASK, false, 0u, 0u};
                int n_ctrl = 0;
                for (int k = 1; k <= 2; ++k) {
                    const unsigned r = refs[k];
                    if (r & (REF_CONST | REF_WIDE)) return false;
                    const int idx = (int)(r & REF_MASK);
                    if (idx >= in.ctrl_slot && idx < in.ctrl_slot + in.nce) { v.c = (unsigned)(idx - in.ctrl_slot); ++n_ctrl; }
                    else {
                        for (const ElemVal& o : ev)
                            if ((int)o.dst == idx) return false;       // product of elem values: not a single term
                        if (idx >= g.x_slot) return false;
                        v.has_a = true; v.a = (unsigned)idx;
                    }
                }
                if (n_ctrl != 1) return false;
                ev.push_back(v);
            }
        }
        ChainDesc& cd = chain;
        cd = ChainDesc{};
        if (!in.par_refs || !in.g_refs || in.n_par > 16 || in.ng > 8) return false;
        cd.n_par = (unsigned)in.n_par; cd.ng = (unsigned)in.ng;
        for (int i = 0; i < in.n_par; ++i) cd.par_off[i] = (uint32_t)(in.par_refs[i] & REF_MASK) * usp;
        for (int k = 0; k < in.ng; ++k) {
            if (in.g_refs[k] & (REF_CONST | REF_WIDE)) return false;  // a constant constraint: leave it to the generic path
            cd.g_off[k] = (uint32_t)(in.g_refs[k] & REF_MASK) * usp;
        }
        if (in.nz > 16 || (in.nz > 0 && !in.z_refs)) return false;
        cd.nz = (unsigned)in.nz;
        for (int zi = 0; zi < in.nz; ++zi) {
            uint32_t is_const = 0;
            if (!source_ref(in, in.z_refs[zi], cd.z_off[zi], is_const)) return false;
            cd.z_const |= is_const << zi;
        }
        int pos = 0, scratch = g.x_slot;                               // units [x_slot, n_units) are dead during the loop
        for (int b = 0; b < in.n_blocks; ++b) {
            const PolyBlock& pb = in.poly[b];
            const uint64_t* code = in.pt + pos;
            const int n = in.blk_npt[b];
            pos += n;
            std::vector<char> used((size_t)n, 0);
            Term t[3];
            for (int k = 0; k < 3; ++k) {
                if (!pb.present[k]) continue;
                const unsigned ref = (pb.off[k] / usp) | (pb.wide[k] ? REF_WIDE : 0u);
                if (pb.wide[k]) {
                    if (k == 2 || !wide_term(in, code, n, ref, b, invariant, ev, used, t[k])) return false;
                } else {
                    if (!narrow_factor(in, ref, invariant, ev, t[k])) return false;
                    t[k].present = true;
                }
            }
            for (int i = 0; i < n; ++i)
                if (!used[i]) return false;                            // a point op the terms do not account for
            uint32_t kind = in.blk_lin[b] ? CHAIN_LINEAR : CHAIN_NEWTON, minv = 0, flops = in.blk_flops[b];
            if (in.blk_lin[b] && (!t[1].present || (t[1].p == 0 && !t[1].has_ctrl)) && scratch + 9 <= in.n_units) {
                kind = CHAIN_LINCONST; minv = (uint32_t)scratch * usp; scratch += 9; flops = CHAIN_LINCONST_FLOPS;
            }
            const uint32_t w[12] = {g.blk_x_o[b][0][0], flops, kind, minv, t[0].n_off, t[0].meta(), t[1].n_off, t[1].meta(),
                                    t[2].n_off, t[2].meta(), 0u, 0u};
            ChainBlockD& bd = cd.blk[b];
            bd.x_off = w[0]; bd.flops = flops; bd.kind = kind; bd.minv_off = minv;
            for (int k = 0; k < 3; ++k) { bd.c[k].n_off = t[k].n_off; bd.c[k].meta = t[k].meta(); }
            const int st = (int)(g.blk_x_o[b][0][0] / usp - (uint32_t)g.x_slot) / 3;      // the block's state
            if (!in.x0_refs || st < 0 || st >= g.ns || !source_ref(in, in.x0_refs[st], bd.x0_off, bd.x0_const)) return false;
            bd.xe_off = (uint32_t)(in.xe_slot + st) * usp;
        }
        // quadrature integrands
        {
            const uint64_t* code = in.pt + pos;
            const int n = in.q_npt;
            std::vector<char> used((size_t)n, 0);
            int nq = 0;
            for (int q = 0; q < in.nq; ++q) {
                const unsigned r = in.fq_refs[q];
                if (r == REF_NONE) return false;                       // a constant quadrature state: generic path
                Term t;
                if (r & REF_CONST) return false;
                if (r & REF_WIDE) {
                    if (!wide_term(in, code, n, r, in.n_blocks, invariant, ev, used, t)) return false;
                } else {
                    if (!narrow_factor(in, r, invariant, ev, t)) return false;
                    t.present = true;
                }
                if (++nq > CHAIN_MAXQ) return false;
                const uint32_t w[4] = {(uint32_t)(in.qe_slot + q) * usp, t.n_off, t.meta(), 0u};
                ChainQuadD& qd = cd.quad[nq - 1];
                qd.q_off = w[0]; qd.n_off = t.n_off; qd.meta = t.meta();
                if (!in.q0_refs || !source_ref(in, in.q0_refs[q], qd.q0_off, qd.q0_const)) return false;
            }
            for (int i = 0; i < n; ++i)
                if (!used[i]) return false;
            if (g.max_it < 1 || g.max_it > 0xFFFF) return false;
            cd.nb = (unsigned)in.n_blocks; cd.nq = (unsigned)nq; cd.max_it = (unsigned)g.max_it;
        }
        nb_out = in.n_blocks;
        return true;
    }

    // The PCHAIN instruction (head, 12 words per block, 4 per quadrature) as the descriptor stands
    void chain_words(std::vector<uint32_t>& out) const
    {
        const ChainDesc& cd = chain;
        const uint32_t head[8] = {X_PCHAIN, cd.nb | (cd.nq << 8), cd.max_it, 0u, 0u, 0u, 0u, 0u};
        out.assign(head, head + 8);
        for (unsigned b = 0; b < cd.nb; ++b) {
            const ChainBlockD& bd = cd.blk[b];
            const uint32_t w[12] = {bd.x_off, bd.flops, bd.kind, bd.minv_off, bd.c[0].n_off, bd.c[0].meta, bd.c[1].n_off, bd.c[1].meta,
                                    bd.c[2].n_off, bd.c[2].meta, 0u, 0u};
            out.insert(out.end(), w, w + 12);
        }
        for (unsigned q = 0; q < cd.nq; ++q) {
            const uint32_t w[4] = {cd.quad[q].q_off, cd.quad[q].n_off, cd.quad[q].meta, 0u};
            out.insert(out.end(), w, w + 4);
        }
    }

    // Liveness-based renumbering of the register-file units of a chain program.  The layout the lowering chose
    // (pyds_b200/tape.py) serves the interpreted element loop: a persistent narrow region plus three units per wide value.
    // A chain program is straight-line code around ONE instruction that keeps the states in registers, so a unit is live
    // from the op that writes a value to the last op that reads it and interval colouring is exact: parameters (written at
    // the tile start), prologue temporaries, coefficients, the scratch of a constant-decay block (9 consecutive units, live
    // inside PCHAIN only), end values and g temporaries share units.  Times: op i reads at 2i and writes at 2i + 1, so an
    // op may overwrite an operand it reads last.  PCHAIN stores its end values after the element loop: they may take the
    // units of the scratch (last read in the loop), not those of its inputs.
    bool compact_units(const AsmInput& in, std::vector<uint32_t>& pre, std::vector<uint32_t>& post)
    {
        const uint32_t usp = US * (uint32_t)in.spt;
        const int n_pre = (int)(pre.size() / 8), n_post = (int)(post.size() / 8);
        if ((int)nread_.size() != n_pre + n_post) return fail("internal: operand counts of the scalar ops");
        const int t_chain = n_pre, t_end = n_pre + 1 + n_post;         // op indices of PCHAIN and END (parameters: op -1)
        struct Val { int start, end, unit; };
        std::vector<Val> vals;
        std::vector<int> cur((size_t)in.n_units + 16, -1);             // value that currently sits in an (old) unit
        struct Patch { uint32_t* where; int val; };
        std::vector<Patch> patches;
        auto old_unit = [&](uint32_t off) -> int { return (int)(off / usp); };
        auto def = [&](uint32_t* where, int t) {
            const int u = old_unit(*where);
            if (u < 0 || u >= (int)cur.size()) return false;
            vals.push_back(Val{2 * t + 1, 2 * t + 1, -1});
            cur[(size_t)u] = (int)vals.size() - 1;
            patches.push_back(Patch{where, cur[(size_t)u]});
            return true;
        };
        auto use = [&](uint32_t* where, int t, int late) {
            const int u = old_unit(*where);
            if (u < 0 || u >= (int)cur.size()) return false;
            if (cur[(size_t)u] < 0) {                                  // read before any write: lives from the tile start
                vals.push_back(Val{-1, -1, -1});
                cur[(size_t)u] = (int)vals.size() - 1;
            }
            Val& v = vals[(size_t)cur[(size_t)u]];
            v.end = std::max(v.end, 2 * t + late);
            patches.push_back(Patch{where, cur[(size_t)u]});
            return true;
        };
        ChainDesc& cd = chain;
        bool ok = true;
        for (unsigned i = 0; i < cd.n_par; ++i) ok = ok && def(&cd.par_off[i], -1);
        // control values read through the descriptor (initial / fixed values): by LDCTRL before the chain when every
        // element sees the same entry, else by the chain itself (any entry, any element)
        auto z_use = [&](unsigned zi, int t, int late) {
            if (zi < cd.nz && !((cd.z_const >> zi) & 1u)) ok = ok && use(&cd.z_off[zi], t, late);
        };
        auto scalar_section = [&](std::vector<uint32_t>& w, int n, int t0, int op0) {
            for (int i = 0; i < n && ok; ++i) {
                uint32_t* o = &w[(size_t)i * 8];
                if (o[0] == (uint32_t)X_LDCTRL) z_use(in.cmap[o[6]], t0 + i, 0);
                const unsigned fl = o[5];
                const int nr = nread_[(size_t)(op0 + i)];
                for (int k = 0; k < 3; ++k) {
                    uint32_t* f = &o[2 + k];
                    if (k < nr) { if (!((fl >> k) & 1u)) ok = ok && use(f, t0 + i, 0); }
                    else if (!((fl >> k) & 1u)) *f = 0u;               // never read for its value: any valid address
                }
                ok = ok && def(&o[1], t0 + i);
            }
        };
        scalar_section(pre, n_pre, 0, 0);
        bool chain_reads_ctrl = false;
        for (unsigned b = 0; b < cd.nb; ++b)
            for (int k = 0; k < 3; ++k) chain_reads_ctrl = chain_reads_ctrl || (cd.blk[b].c[k].meta & CT_HAS_CTRL);
        for (unsigned q = 0; q < cd.nq; ++q) chain_reads_ctrl = chain_reads_ctrl || (cd.quad[q].meta & CT_HAS_CTRL);
        for (int e = 0; chain_reads_ctrl && e < in.nfe; ++e)
            for (int c = 0; c < in.nce; ++c) z_use(in.cmap[e * in.nce + c], t_chain, 1);
        // PCHAIN: inputs (late = 1: still held while the outputs are written), outputs; the scratch further down
        for (unsigned b = 0; b < cd.nb && ok; ++b) {
            ChainBlockD& bd = cd.blk[b];
            for (int k = 0; k < 3; ++k)
                if ((bd.c[k].meta & CT_PRESENT) && (bd.c[k].meta & CT_HAS_N)) ok = ok && use(&bd.c[k].n_off, t_chain, 1);
            if (!bd.x0_const) ok = ok && use(&bd.x0_off, t_chain, 1);
        }
        for (unsigned q = 0; q < cd.nq && ok; ++q) {
            ChainQuadD& qd = cd.quad[q];
            if ((qd.meta & CT_PRESENT) && (qd.meta & CT_HAS_N)) ok = ok && use(&qd.n_off, t_chain, 1);
            if (!qd.q0_const) ok = ok && use(&qd.q0_off, t_chain, 1);
        }
        for (unsigned b = 0; b < cd.nb && ok; ++b) ok = ok && def(&cd.blk[b].xe_off, t_chain);
        for (unsigned q = 0; q < cd.nq && ok; ++q) ok = ok && def(&cd.quad[q].q_off, t_chain);
        scalar_section(post, n_post, t_chain + 1, n_pre);
        for (unsigned k = 0; k < cd.ng && ok; ++k) ok = ok && use(&cd.g_off[k], t_end, 0);
        if (!ok) return fail("internal: unit out of range while compacting");

        // interval colouring: values that are live when PCHAIN starts first (they all overlap there), then the scratch blocks
        // right above them, then everything else in order of its start -- lowest unit whose intervals do not overlap
        std::vector<std::vector<std::pair<int, int>>> busy;
        auto fits = [&](int u, int s, int e) {
            if (u >= (int)busy.size()) return true;
            for (const auto& iv : busy[(size_t)u])
                if (s <= iv.second && iv.first <= e) return false;
            return true;
        };
        auto take = [&](int u, int s, int e) {
            if (u >= (int)busy.size()) busy.resize((size_t)u + 1);
            busy[(size_t)u].push_back(std::make_pair(s, e));
        };
        auto place = [&](Val& v) {
            int u = 0;
            while (!fits(u, v.start, v.end)) ++u;
            take(u, v.start, v.end);
            v.unit = u;
        };
        const int tc0 = 2 * t_chain, tc1 = 2 * t_chain + 1;
        std::vector<int> order;
        for (size_t i = 0; i < vals.size(); ++i)
            if (vals[i].start <= tc0 && vals[i].end >= tc0) order.push_back((int)i);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return vals[(size_t)a].start < vals[(size_t)b].start; });
        for (int i : order) place(vals[(size_t)i]);
        for (unsigned b = 0; b < cd.nb; ++b) {
            if (cd.blk[b].kind != CHAIN_LINCONST) continue;
            int u = 0;
            for (;; ++u) {
                bool free9 = true;
                for (int k = 0; k < 9 && free9; ++k) free9 = fits(u + k, tc0, tc0);
                if (free9) break;
            }
            for (int k = 0; k < 9; ++k) take(u + k, tc0, tc0);
            cd.blk[b].minv_off = (uint32_t)u * usp;
        }
        order.clear();
        for (size_t i = 0; i < vals.size(); ++i)
            if (vals[i].unit < 0) order.push_back((int)i);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return vals[(size_t)a].start < vals[(size_t)b].start; });
        for (int i : order) place(vals[(size_t)i]);
        for (const Patch& p : patches) *p.where = (uint32_t)vals[(size_t)p.val].unit * usp;
        chain_units = (int)busy.size();
        if (chain_units > in.n_units) return fail("internal: compaction grew the register file");
        return true;
    }

    bool fail(const std::string& m)
    {
        err = m;
        return false;
    }
    // Is any of the wide value's three register-file units [unit, unit + 2] read by anything but the two point-tape ops
    // at skip, skip + 1 (the SQR / MUL pair that produces it) and coefficient 0 of block own_block?  Readers: every
    // point-tape op (all blocks + the quadrature tape), the quadrature integrand table, the other polynomial
    // coefficients and the f / J tables of the generic blocks.
    bool reads_units_elsewhere(const AsmInput& in, unsigned unit, int skip, int own_block) const
    {
        auto hit = [&](unsigned ref) { return !(ref & REF_CONST) && (ref & REF_MASK) + ((ref & REF_WIDE) ? 2u : 0u) >= unit && (ref & REF_MASK) <= unit + 2u; };
        int total = in.q_npt;
        for (int b = 0; b < in.n_blocks; ++b) total += in.blk_npt[b];
        for (int j = 0; j < total; ++j) {
            if (j == skip || j == skip + 1) continue;
            unsigned op, refs[4];
            split(in.pt[j], op, refs);
            for (int k = 1; k <= arity(op); ++k)
                if (hit(refs[k])) return true;
        }
        for (int q = 0; q < in.nq; ++q)
            if (in.fq_refs[q] != REF_NONE && hit(in.fq_refs[q])) return true;
        const uint32_t usp = US * (uint32_t)in.spt;
        auto hit_off = [&](uint32_t off, bool wide) { return off / usp + (wide ? 2u : 0u) >= unit && off / usp <= unit + 2u; };
        for (int b = 0; b < in.n_blocks; ++b) {
            if (in.poly[b].is_poly) {
                for (int k = 0; k < 3; ++k)
                    if (in.poly[b].present[k] && !(b == own_block && k == 0) && hit_off(in.poly[b].off[k], in.poly[b].wide[k])) return true;
            } else {
                const ExecImage& g = *in.img;
                for (int n = 0; n < in.blk_ns[b]; ++n)
                    for (int p3 = 0; p3 < 3; ++p3)
                        if (hit_off(g.blk_f_o[b][n][p3], false)) return true;
                for (int i = 0; i < in.blk_ns[b] * in.blk_ns[b]; ++i)
                    if ((g.blk_jmask[b] >> i) & 1u)
                        for (int p3 = 0; p3 < 3; ++p3)
                            if (hit_off(g.blk_j_o[b][i][p3], false)) return true;
            }
        }
        return false;
    }

    void emit(uint32_t x, uint32_t y = 0, uint32_t z = 0, uint32_t w = 0, uint32_t x1 = 0, uint32_t y1 = 0, uint32_t z1 = 0,
              uint32_t w1 = 0)
    {
        const uint32_t v[8] = {x, y, z, w, x1, y1, z1, w1};
        words.insert(words.end(), v, v + 8);
    }
    uint32_t slot_off(unsigned ref) const { return (ref & REF_MASK) * US * (uint32_t)in_->spt; }
    bool check_slot(unsigned ref)
    {
        if ((int)(ref & REF_MASK) + ((ref & REF_WIDE) ? 2 : 0) >= in_->n_units) return fail("slot out of range");
        return true;
    }
    static int arity(unsigned op)
    {
        if (op == OP_FMA || op == OP_FMS || op == OP_FNMA) return 3;
        if (op == OP_ADD || op == OP_SUB || op == OP_MUL || op == OP_DIV || op == OP_POW || op == OP_MIN || op == OP_MAX)
            return 2;
        return 1;
    }
    static void split(uint64_t w, unsigned& op, unsigned (&refs)[4])
    {
        op = (unsigned)(w & 0xFF);
        refs[0] = (unsigned)(w >> 8) & 0x3FFFu;
        refs[1] = (unsigned)(w >> 22) & 0x3FFFu;
        refs[2] = (unsigned)(w >> 36) & 0x3FFFu;
        refs[3] = (unsigned)(w >> 50) & 0x3FFFu;
    }
    static bool is_cold(unsigned op)
    {
        return op == OP_EXP || op == OP_LOG || op == OP_SQRT || op == OP_SIN || op == OP_COS || op == OP_TANH || op == OP_POW;
    }

    // pro / elem / g tapes: one value per scenario; operands are slots or constant-pool entries
    bool scalar_op(uint64_t w)
    {
        unsigned op, refs[4];
        split(w, op, refs);
        if (op >= OP_BC3) return fail("unknown opcode in a scalar tape");
        static const int map[OP_BC3] = {X_MOV1, X_NEG1, X_ADD1, X_SUB1, X_MUL1, X_DIV1, X_FMA1, X_FMS1, X_FNMA1, X_SQR1,
                                        X_RCP1, X_COLD1, X_COLD1, X_COLD1, X_COLD1, X_COLD1, X_COLD1, X_COLD1, X_ABS1,
                                        X_MIN1, X_MAX1};
        const int n = arity(op);
        if (refs[0] & REF_CONST) return fail("scalar tape writes a constant");
        if (!check_slot(refs[0] & ~REF_WIDE)) return false;
        uint32_t off[3] = {0, 0, 0}, flags = 0;
        // constant operands: relative to cpool - 1 (cpool[-1] = 0.0); chain programs: SPT-wide entries after the three
        // specials 1.0, -0.0, 0.0
        const uint32_t cw = chain_mode_ ? 8u * (uint32_t)in_->spt : 8u, cpre = chain_mode_ ? 3u : 1u;
        for (int k = 0; k < n; ++k) {
            const unsigned r = refs[k + 1];
            if (r & REF_CONST) {
                flags |= 1u << k;
                off[k] = (r == REF_NONE) ? (cpre - 1u) * cw : ((r & REF_MASK) + cpre) * cw;
            } else {
                if (!check_slot(r & ~REF_WIDE)) return false;
                off[k] = slot_off(r);
            }
        }
        if (chain_mode_) {
            // every FMA-class op is FMAG: r = fma(+-a, b, +-c) with 1.0 / -0.0 from the pool (bit-identical results)
            const uint32_t ONE = 0u, NEGZERO = cw;
            struct Opnd { uint32_t off; bool cst; };
            const Opnd A{off[0], (flags & 1u) != 0}, B{off[1], (flags & 2u) != 0}, C{off[2], (flags & 4u) != 0};
            const Opnd one{ONE, true}, nz{NEGZERO, true};
            Opnd a = A, b = one, c = nz;
            bool neg_a = false, neg_c = false, fmag = true;
            switch (op) {
            case OP_MOV: break;
            case OP_NEG: neg_a = true; break;
            case OP_ADD: c = B; break;
            case OP_SUB: a = B; neg_a = true; c = A; break;
            case OP_MUL: b = B; break;
            case OP_SQR: b = A; break;
            case OP_FMA: b = B; c = C; break;
            case OP_FMS: b = B; c = C; neg_c = true; break;
            case OP_FNMA: b = B; c = C; neg_a = true; break;
            default: fmag = false; break;
            }
            if (fmag) {
                const uint32_t f = (a.cst ? 1u : 0u) | (b.cst ? 2u : 0u) | (c.cst ? 4u : 0u) | (neg_a ? 8u : 0u) | (neg_c ? 16u : 0u);
                emit(X_FMAG, slot_off(refs[0]), a.off, b.off, c.off, f, 0u);
                nread_.push_back(3);
                return true;
            }
        }
        if (n == 1) {                                     // cold unary ops are dispatched as binary: b = a
            off[1] = off[0];
            if (flags & 1u) flags |= 2u;
        }
        emit((uint32_t)map[op], slot_off(refs[0]), off[0], off[1], off[2], flags, is_cold(op) ? op : 0u);
        if (chain_mode_) nread_.push_back((unsigned char)(n == 1 ? 2 : n));
        return true;
    }

    // One point tape.  Peephole: `t = x^2` followed by `d = n * t` with t read by nothing else (not later in the tape, not by
    // the op that consumes the tape's outputs: `live`) becomes one MULSQ op (second-order mass-action rates k * c^2).
    bool point_tape(const uint64_t* code, int n, const std::vector<unsigned>& live)
    {
        for (int i = 0; i < n; ++i) {
            unsigned op, refs[4];
            split(code[i], op, refs);
            if (op == OP_SQR && i + 1 < n && (refs[0] & REF_WIDE) && (refs[1] & REF_WIDE)) {
                unsigned op2, r2[4];
                split(code[i + 1], op2, r2);
                const unsigned t = refs[0] & REF_MASK;
                unsigned na = r2[1], wb = r2[2];
                if ((na & REF_WIDE) && !(wb & REF_WIDE)) { const unsigned s = na; na = wb; wb = s; }
                bool fuse = op2 == OP_MUL && (r2[0] & REF_WIDE) && !(na & (REF_WIDE | REF_CONST)) && (wb & REF_WIDE) &&
                            (wb & REF_MASK) == t;
                for (size_t k = 0; fuse && k < live.size(); ++k) fuse = live[k] != t || (r2[0] & REF_MASK) == t;
                for (int j = i + 2; fuse && j < n; ++j) {
                    unsigned opj, rj[4];
                    split(code[j], opj, rj);
                    for (int k = 1; k <= arity(opj); ++k) fuse = fuse && ((rj[k] & REF_CONST) || (rj[k] & REF_MASK) != t);
                    if ((rj[0] & REF_MASK) == t) break;              // overwritten before any later read
                }
                if (fuse) {
                    if (!check_slot(r2[0]) || !check_slot(na) || !check_slot(refs[1])) return false;
                    emit(X_MULSQ_NW, slot_off(r2[0]), slot_off(na), slot_off(refs[1]));
                    ++i;
                    continue;
                }
            }
            if (!point_op(code[i])) return false;
        }
        return true;
    }

    // point tapes: three collocation points per instruction; operands are narrow or wide slots
    bool point_op(uint64_t w)
    {
        unsigned op, refs[4];
        split(w, op, refs);
        if (op >= OP_BC3) return fail("opcode not allowed in the point tape");
        const int n = arity(op);
        for (int k = 0; k <= n; ++k) {
            if (refs[k] & REF_CONST) return fail("constant operand in the point tape");
            if (!check_slot(refs[k])) return false;
        }
        if (!(refs[0] & REF_WIDE)) return fail("point tape writes a narrow slot");
        unsigned a = refs[1], b = refs[2], c = refs[3];
        auto wide = [](unsigned r) { return (r & REF_WIDE) != 0; };
        const uint32_t d = slot_off(refs[0]);
        if (is_cold(op)) {
            if (n == 1) b = a;
            emit(X_COLD_W, d, slot_off(a), slot_off(b), 0u, (wide(a) ? 1u : 0u) | (wide(b) ? 2u : 0u), op);
            return true;
        }
        switch (op) {
        case OP_FMA: case OP_FMS: case OP_FNMA: {
            if (wide(a) && !wide(b)) { const unsigned t = a; a = b; b = t; }
            if (!wide(b)) return fail("point-level product of two narrow values (should have been hoisted)");
            static const uint32_t tab[3][4] = {{X_FMA_NWN, X_FMA_NWW, X_FMA_WWN, X_FMA_WWW},
                                               {X_FMS_NWN, X_FMS_NWW, X_FMS_WWN, X_FMS_WWW},
                                               {X_FNMA_NWN, X_FNMA_NWW, X_FNMA_WWN, X_FNMA_WWW}};
            const int fam = op == OP_FMA ? 0 : (op == OP_FMS ? 1 : 2);
            emit(tab[fam][(wide(a) ? 2 : 0) + (wide(c) ? 1 : 0)], d, slot_off(a), slot_off(b), slot_off(c));
            return true;
        }
        case OP_MUL: case OP_ADD: case OP_MIN: case OP_MAX: {
            if (wide(a) && !wide(b)) { const unsigned t = a; a = b; b = t; }
            if (!wide(b)) return fail("point-level binary op on two narrow values (should have been hoisted)");
            static const uint32_t tab[4][2] = {{X_MUL_NW, X_MUL_WW}, {X_ADD_NW, X_ADD_WW}, {X_MIN_NW, X_MIN_WW}, {X_MAX_NW, X_MAX_WW}};
            const int fam = op == OP_MUL ? 0 : (op == OP_ADD ? 1 : (op == OP_MIN ? 2 : 3));
            emit(tab[fam][wide(a) ? 1 : 0], d, slot_off(a), slot_off(b));
            return true;
        }
        case OP_SUB: case OP_DIV: {
            if (!wide(a) && !wide(b)) return fail("point-level binary op on two narrow values (should have been hoisted)");
            static const uint32_t tab[2][3] = {{X_SUB_NW, X_SUB_WN, X_SUB_WW}, {X_DIV_NW, X_DIV_WN, X_DIV_WW}};
            emit(tab[op == OP_SUB ? 0 : 1][!wide(a) ? 0 : (!wide(b) ? 1 : 2)], d, slot_off(a), slot_off(b));
            return true;
        }
        case OP_MOV:
            emit(wide(a) ? X_MOV_W : X_BC_N, d, slot_off(a));
            return true;
        case OP_SQR: case OP_NEG: case OP_RCP: case OP_ABS: {
            if (!wide(a)) return fail("point-level unary op on a narrow value (should have been hoisted)");
            const uint32_t code = op == OP_SQR ? X_SQR_W : (op == OP_NEG ? X_NEG_W : (op == OP_RCP ? X_RCP_W : X_ABS_W));
            emit(code, d, slot_off(a));
            return true;
        }
        default:
            return fail("opcode not allowed in the point tape");
        }
    }
};

}  // namespace pydsb
