// C-ABI of libpyds_b200.so (see include/pydsb.h).  Host side only: blob parsing, device images,
// launch configuration, the host loop of the recourse optimiser, pinned staging for the *_host
// entry points.  No CPU compute fallback: without a CUDA device every entry point fails with
// PYDSB_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <algorithm>
#include <utility>
#include <vector>

#include "../../include/pydsb.h"
#include "assembler.h"
#include "chain_shapes.h"
#include "recourse_kernel.cuh"
#include "recourse_chain.cuh"

using namespace pydsb;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
    g_err = msg;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return fail(PYDSB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
    } while (0)

struct Section {
    char tag[5];
    uint32_t dtype, count;
    const unsigned char* data;
};
typedef std::vector<Section> Sections;

const Section* find(const Sections& secs, const char* tag)
{
    for (const auto& s : secs)
        if (strncmp(s.tag, tag, 4) == 0) return &s;
    return nullptr;
}

// One lowered program (0: feasibility, 1: feasibility + control sensitivities) on the device.
struct Program {
    ExecImage img{};
    void* d_consts = nullptr;
    void* d_tabs = nullptr;
    std::vector<uint32_t> xprog;        // the assembled program (program 0: whole scenario; program 1: tape segments)
    std::vector<uint32_t> xprog_interp; // program 0 with the element loop interpreted (pydsb_state_solve), when xprog is fused
    int chain_nb = 0;                   // blocks of the fused element loop (X_PCHAIN) in xprog; 0: none
    int chain_units = 0;                // register-file units xprog touches after the assembler's compaction (0: img.n_units)
    int chain_smem_bytes = 0;           // shared memory per CTA of the fused program (xprog_interp needs img.smem_bytes)
    ChainDesc chain{};                  // its descriptor, copied to the constant bank (c_chain) before a launch
    unsigned long long shape_sb = 0;    // compile-time shape words of the chain (poly_chain.cuh, chain_shape_bits)
    unsigned shape_sq = 0;
    RChainDesc rchain{};                // program 2: the chain with control sensitivities (recourse_chain.cuh)
    PolyBlock poly[MAX_BLOCKS];         // program 0: blocks lowered to polynomial coefficients
    bool present = false;
};

struct Model {
    int device = 0;
    Program prog[3];                    // 0 feasibility, 1 coupled-block sensitivities, 2 chain + dg (sensitivities in registers)
    std::vector<double> z_lo, z_hi;
    std::vector<unsigned char> z_fix;
    double* d_zlo = nullptr;
    double* d_zhi = nullptr;
    unsigned char* d_zfix = nullptr;
    unsigned long long* d_counters = nullptr;
    double* d_partial = nullptr;
    size_t partial_cap = 0;
    int num_sms = 0, blocks_per_sm = 0, regs = 0;
    unsigned long long launches = 0;
    double* h_pin = nullptr;            // pinned + device staging for the *_host entry points
    size_t h_pin_cap = 0;
    double* d_stage = nullptr;
    cudaStream_t own_stream = nullptr;
    // recourse workspace kept between calls while it is small (the nested-sampling loop calls with a few hundred
    // groups, thousands of times): cudaMalloc/cudaFree cost more than the kernels there
    void* rws[3] = {nullptr, nullptr, nullptr};
    size_t rws_cap[3] = {0, 0, 0};
    // eval_host works on own_stream while the device entry points work on the caller's stream, and both use the model's
    // partial-sum buffer and the program window in constant memory: own_stream waits for the last asynchronous launch
    cudaEvent_t ev_async = nullptr;
    bool async_pending = false;
    // per-kernel device time (pydsb_set_profiling): event pairs around launches, folded into ms[] by prof_collect
    bool prof_on = false;
    struct ProfRec { int kind; cudaEvent_t e0, e1; };
    std::vector<ProfRec> prof_recs;
    double prof_ms[3] = {0.0, 0.0, 0.0};
    unsigned long long prof_n[3] = {0, 0, 0};
};

// RAII: events around one launch of kernel class `kind` (0 feas, 1 sens, 2 lm_update) when profiling is on
struct ProfScope {
    Model* m; cudaStream_t st; int kind; cudaEvent_t e0 = nullptr, e1 = nullptr;
    ProfScope(Model* m_, int kind_, cudaStream_t st_) : m(m_), st(st_), kind(kind_)
    {
        if (!m->prof_on) return;
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e0 = e1 = nullptr; return; }
        cudaEventRecord(e0, st);
    }
    ~ProfScope()
    {
        if (!e0 || !e1) return;
        cudaEventRecord(e1, st);
        m->prof_recs.push_back(Model::ProfRec{kind, e0, e1});
    }
};

void prof_collect(Model* m)
{
    for (const Model::ProfRec& r : m->prof_recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.e1) == cudaSuccess && cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            m->prof_ms[r.kind] += ms;
            m->prof_n[r.kind] += 1;
        }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    m->prof_recs.clear();
}

typedef void (*FeasFn)(const ExecImage, const LaunchArgs);
typedef void (*SensFn)(const ExecImage, const RecourseArgs);

// MAXNS = largest coupled block, SPT = scenarios per thread.  SPT = 2: both scenarios of a thread are moved by
// one LDS.128 and share every fetch, dispatch and address of the interpreter (~30 % fewer instructions per
// scenario) at half the resident warps.  Measured on C3: 20 % slower than SPT = 1 while the Newton loops ran through
// the dispatch loop (bound by the dependent fetch -> dispatch -> load -> math chain, i.e. by warps in flight),
// 7 % faster once the one-state blocks run in registers (POLY1; bound by instruction issue).  Default: 2 when
// every block is a polynomial one-state block, else 1; PYDSB_SPT=1|2 overrides (2 needs one-state blocks: the two
// Newton systems have to fit the register file).
struct ChainKernel {
    unsigned long long sb;
    unsigned sq;
    FeasFn plain, pack;
};
#define PYDSB_X(SB_, SQ_, NB_) {SB_, SQ_, feas_kernel<1, 2, false, false, NB_, SB_, SQ_>, feas_kernel<1, 2, false, true, NB_, SB_, SQ_>},
const ChainKernel g_chain_kernels[] = {PYDSB_CHAIN_SHAPES(PYDSB_X)};
#undef PYDSB_X

// instantiation with the chain's shape folded at compile time, if that shape is registered
const ChainKernel* chain_kernel_for(const Program& p, int spt)
{
    if (p.chain_nb <= 0 || spt != 2 || getenv("PYDSB_GENERIC_CHAIN")) return nullptr;
    for (const ChainKernel& k : g_chain_kernels)
        if (k.sb == p.shape_sb && k.sq == p.shape_sq) return &k;
    return nullptr;
}

FeasFn feas_for(int ns, int spt, const Program* p = nullptr)
{
    const int chain = p ? p->chain_nb : 0;
    if (chain > 0 && ns == 1) {              // fused element loop (poly_chain.cuh): 2 or 3 blocks in registers
        if (const ChainKernel* k = chain_kernel_for(*p, spt)) return k->plain;
        if (spt == 2) return chain <= 2 ? feas_kernel<1, 2, false, false, 2> : feas_kernel<1, 2, false, false, 3>;
        return chain <= 2 ? feas_kernel<1, 1, false, false, 2> : feas_kernel<1, 1, false, false, 3>;
    }
    switch (ns * 10 + spt) {
    case 11: return feas_kernel<1, 1>;
    case 12: return feas_kernel<1, 2>;
    case 21: return feas_kernel<2, 1>;
    case 31: return feas_kernel<3, 1>;
    default: return nullptr;
    }
}

// the trajectory-writing variants of the same kernels (pydsb_state_solve)
FeasFn traj_for(int ns, int spt)
{
    switch (ns * 10 + spt) {
    case 11: return feas_kernel<1, 1, true>;
    case 12: return feas_kernel<1, 2, true>;
    case 21: return feas_kernel<2, 1, true>;
    case 31: return feas_kernel<3, 1, true>;
    default: return nullptr;
    }
}

// small theta sets: several design points per tile
FeasFn pack_for(int ns, int spt, const Program* p = nullptr)
{
    const int chain = p ? p->chain_nb : 0;
    if (chain > 0 && ns == 1) {
        if (const ChainKernel* k = chain_kernel_for(*p, spt)) return k->pack;
        if (spt == 2) return chain <= 2 ? feas_kernel<1, 2, false, true, 2> : feas_kernel<1, 2, false, true, 3>;
        return chain <= 2 ? feas_kernel<1, 1, false, true, 2> : feas_kernel<1, 1, false, true, 3>;
    }
    switch (ns * 10 + spt) {
    case 11: return feas_kernel<1, 1, false, true>;
    case 12: return feas_kernel<1, 2, false, true>;
    case 21: return feas_kernel<2, 1, false, true>;
    case 31: return feas_kernel<3, 1, false, true>;
    default: return nullptr;
    }
}

int image_maxns(const ExecImage& g)
{
    int maxns = 1;
    for (int b = 0; b < g.n_blocks; ++b)
        if (g.blk_ns[b] > maxns) maxns = g.blk_ns[b];
    return maxns;
}

// scenarios per thread of the feasibility program; PYDSB_SPT=1|2 overrides (experiments)
int choose_spt(int maxns, bool all_poly)
{
    int spt = (maxns == 1 && all_poly) ? 2 : 1;
    if (const char* e = getenv("PYDSB_SPT")) {
        const int v = atoi(e);
        if (v == 1 || (v == 2 && maxns == 1)) spt = v;
    }
    return spt;
}
SensFn sens_for(int ns, bool tri)
{
    switch (ns * 2 + (tri && ns > 1 ? 1 : 0)) {
    case 2: return sens_kernel<1, false>;
    case 4: return sens_kernel<2, false>;
    case 5: return sens_kernel<2, true>;
    case 6: return sens_kernel<3, false>;
    case 7: return sens_kernel<3, true>;
    default: return nullptr;
    }
}

// the chain evaluator of the control optimisation: the instantiation with the chain's shape folded at compile time when
// that shape is registered (chain_shapes.h), else the generic one
#define PYDSB_X(SB_, SQ_, NB_) {SB_, SQ_, rsens_kernel<NB_, SB_, SQ_>},
const struct { unsigned long long sb; unsigned sq; SensFn fn; } g_rsens_kernels[] = {PYDSB_CHAIN_SHAPES(PYDSB_X)};
#undef PYDSB_X
SensFn rsens_for(const Program& p)
{
    if (!getenv("PYDSB_GENERIC_CHAIN"))
        for (const auto& k : g_rsens_kernels)
            if (k.sb == p.shape_sb && k.sq == p.shape_sq) return k.fn;
    return p.chain_nb <= 2 ? rsens_kernel<2> : rsens_kernel<3>;
}

// the whole shared-control optimisation of a group in one launch (theta set within one tile)
typedef void (*RFusedFn)(const ExecImage, const RecourseArgs, const LmArgs);
#define PYDSB_X(SB_, SQ_, NB_) {SB_, SQ_, rfused_kernel<NB_, SB_, SQ_>},
const struct { unsigned long long sb; unsigned sq; RFusedFn fn; } g_rfused_kernels[] = {PYDSB_CHAIN_SHAPES(PYDSB_X)};
#undef PYDSB_X
RFusedFn rfused_for(const Program& p)
{
    if (!getenv("PYDSB_GENERIC_CHAIN"))
        for (const auto& k : g_rfused_kernels)
            if (k.sb == p.shape_sb && k.sq == p.shape_sq) return k.fn;
    return p.chain_nb <= 2 ? rfused_kernel<2> : rfused_kernel<3>;
}

int align_up(int v, int a) { return (v + a - 1) / a * a; }

void free_program(Program& p)
{
    if (p.d_consts) cudaFree(p.d_consts);
    if (p.d_tabs) cudaFree(p.d_tabs);
    p.d_consts = p.d_tabs = nullptr;
}

bool resolve_table(const Sections& secs, const char* tag, int n, int cap, unsigned (*off)[3], unsigned* mask, unsigned US)
{
    const Section* s = find(secs, tag);
    if (!s || (int)s->count < n || n > cap) return false;
    const uint16_t* v = reinterpret_cast<const uint16_t*>(s->data);
    *mask = 0;
    for (int i = 0; i < n; ++i) {
        off[i][0] = off[i][1] = off[i][2] = 0;
        if (v[i] == REF_NONE) continue;
        if (v[i] & REF_CONST) return false;
        *mask |= 1u << i;
        for (int p3 = 0; p3 < 3; ++p3)
            off[i][p3] = ((v[i] & REF_MASK) + ((v[i] & REF_WIDE) ? p3 : 0)) * US;
    }
    return true;
}

// Parse one program's sections into an ExecImage + device buffers.
// mode: 0 feasibility program, 1 coupled-block sensitivity program, 2 chain program with dg/dxe, dg/dqe (rsens_kernel)
int build_program(const Sections& secs, int mode, Program& P, bool upload = true)
{
    const bool sens = mode == 1, rchain = mode == 2;
    static const char* need[] = {"HEAD", "DBLS", "CNST", "ADOT", "BWTS", "BIGM", "ZFIX", "TPRO", "TELM", "TPNT", "TGEE",
                                 "RPAR", "RX0_", "RQ0_", "RZ__", "RF__", "RJ__", "RFQ_", "RG__", "CMAP"};
    for (const char* t : need)
        if (!find(secs, t)) return fail(PYDSB_ERR_INVALID, std::string("blob lacks section ") + t);
    const Section* head = find(secs, "HEAD");
    if (head->count < 19) return fail(PYDSB_ERR_INVALID, "HEAD too short");
    const uint32_t* H = reinterpret_cast<const uint32_t*>(head->data);
    ExecImage& g = P.img;
    g.ns = H[0]; g.nq = H[1]; g.ng = H[2]; g.d_dim = H[3]; g.p_dim = H[4]; g.nz = H[5]; g.nfe = H[6];
    const int ncp = H[7];
    g.nce = H[8]; g.n_units = H[9]; g.ctrl_slot = H[10]; g.xe_slot = H[11]; g.qe_slot = H[12];
    g.x_slot = H[13]; g.max_it = H[14]; g.fq_state_dep = H[16]; g.n_blocks = H[18];
    if (ncp != 3) return fail(PYDSB_ERR_UNSUPPORTED, "kernels are built for ncp = 3 (Radau)");
    if (g.ns < 1 || g.ns > MAX_STATES) return fail(PYDSB_ERR_UNSUPPORTED, "1..8 dynamic states are supported");
    if (g.n_blocks < 1 || g.n_blocks > MAX_BLOCKS) return fail(PYDSB_ERR_UNSUPPORTED, "1..4 blocks of coupled states are supported");
    {
        // scenarios per thread: the byte offsets below are laid out for slot stride spt*BLOCK doubles
        const Section* bs0 = find(secs, "BLKS");
        int maxns = 1;
        bool all_poly = bs0 != nullptr;
        if (bs0)
            for (int b = 0; b < g.n_blocks && 5 * b + 1 < (int)bs0->count; ++b) {
                const uint16_t* B0 = reinterpret_cast<const uint16_t*>(bs0->data);
                if (B0[5 * b] > maxns) maxns = B0[5 * b];
                all_poly = all_poly && (B0[5 * b + 1] & 2u);
            }
        g.spt = (sens || rchain) ? 1 : choose_spt(maxns, all_poly);
    }
    const unsigned US = (unsigned)g.spt * BLOCK * 8u;      // bytes between consecutive slots
    {
        // blocks: [nsb, linear, s0, s1, s2] each; tape lengths per block + quadrature tape
        const Section *bs = find(secs, "BLKS"), *bt = find(secs, "BLKT"), *bf = find(secs, "BLKF");
        if (!bs || !bt || !bf || (int)bs->count < 5 * g.n_blocks || (int)bt->count < g.n_blocks + 1 || (int)bf->count < g.n_blocks)
            return fail(PYDSB_ERR_INVALID, "bad block tables");
        const uint16_t* B = reinterpret_cast<const uint16_t*>(bs->data);
        const uint32_t* T = reinterpret_cast<const uint32_t*>(bt->data);
        const uint32_t* F = reinterpret_cast<const uint32_t*>(bf->data);
        const Section *rf = find(secs, "RF__"), *rj = find(secs, "RJ__");
        const uint16_t* RF = reinterpret_cast<const uint16_t*>(rf->data);
        const uint16_t* RJ = reinterpret_cast<const uint16_t*>(rj->data);
        int pc = 0, fpos = 0, jpos = 0, seen = 0;
        for (int b = 0; b < g.n_blocks; ++b) {
            const int nsb = B[5 * b];
            if (nsb < 1 || nsb > 3)
                return fail(PYDSB_ERR_UNSUPPORTED, "register-resident Newton kernels cover 1..3 dynamic states per coupled block");
            const unsigned kind = B[5 * b + 1];            // bit 0: linear, bit 1: polynomial in its own (single) state
            g.blk_ns[b] = nsb; g.blk_lin[b] = kind & 1u; g.blk_pc0[b] = pc; g.blk_npt[b] = (int)T[b]; g.blk_flops[b] = F[b];
            pc += (int)T[b];
            if ((int)rf->count < fpos + nsb || (int)rj->count < jpos + nsb * nsb) return fail(PYDSB_ERR_INVALID, "bad f/J tables");
            g.blk_jmask[b] = 0;
            P.poly[b] = PolyBlock{};
            if (kind & 2u) {
                // RF__/RJ__/RP2_ hold the references of c0/c1/c2:  f = c0 + c1 x + c2 x^2
                const Section* rp = find(secs, "RP2_");
                if (sens || nsb != 1 || !rp || (int)rp->count <= b) return fail(PYDSB_ERR_INVALID, "bad polynomial block");
                const int st = B[5 * b + 2];
                if (st >= g.ns) return fail(PYDSB_ERR_INVALID, "bad state index in block table");
                seen |= 1 << st;
                for (int p3 = 0; p3 < 3; ++p3) g.blk_x_o[b][0][p3] = (unsigned)(g.x_slot + 3 * st + p3) * US;
                const unsigned refs[3] = {RF[fpos], RJ[jpos], reinterpret_cast<const uint16_t*>(rp->data)[b]};
                P.poly[b].is_poly = true;
                for (int k = 0; k < 3; ++k) {
                    if (refs[k] == REF_NONE) continue;
                    if (refs[k] & REF_CONST) return fail(PYDSB_ERR_INVALID, "constant reference in a polynomial block");
                    if ((int)(refs[k] & REF_MASK) + ((refs[k] & REF_WIDE) ? 2 : 0) >= g.n_units) return fail(PYDSB_ERR_INVALID, "slot out of range");
                    P.poly[b].present[k] = true;
                    P.poly[b].wide[k] = (refs[k] & REF_WIDE) != 0;
                    P.poly[b].off[k] = (refs[k] & REF_MASK) * US;
                }
                fpos += 1; jpos += 1;
                continue;
            }
            for (int n = 0; n < nsb; ++n) {
                const int st = B[5 * b + 2 + n];
                if (st >= g.ns) return fail(PYDSB_ERR_INVALID, "bad state index in block table");
                seen |= 1 << st;
                const unsigned rfv = RF[fpos + n];
                if (rfv == REF_NONE || (rfv & REF_CONST)) return fail(PYDSB_ERR_UNSUPPORTED, "a dynamic state has an identically zero right-hand side");
                for (int p3 = 0; p3 < 3; ++p3) {
                    g.blk_x_o[b][n][p3] = (unsigned)(g.x_slot + 3 * st + p3) * US;
                    g.blk_f_o[b][n][p3] = ((rfv & REF_MASK) + ((rfv & REF_WIDE) ? p3 : 0)) * US;
                }
            }
            for (int i = 0; i < nsb * nsb; ++i) {
                const unsigned v = RJ[jpos + i];
                for (int p3 = 0; p3 < 3; ++p3) g.blk_j_o[b][i][p3] = 0;
                if (v == REF_NONE) continue;
                if (v & REF_CONST) return fail(PYDSB_ERR_INVALID, "constant reference in J table");
                g.blk_jmask[b] |= 1u << i;
                for (int p3 = 0; p3 < 3; ++p3) g.blk_j_o[b][i][p3] = ((v & REF_MASK) + ((v & REF_WIDE) ? p3 : 0)) * US;
            }
            fpos += nsb; jpos += nsb * nsb;
        }
        if (seen != (1 << g.ns) - 1) return fail(PYDSB_ERR_INVALID, "blocks do not cover every dynamic state");
        g.q_pc0 = pc; g.q_npt = (int)T[g.n_blocks];
    }
    if (g.ng < 1 || g.ng > 8 || g.nfe < 1 || g.p_dim < 1 || g.d_dim < 0) return fail(PYDSB_ERR_INVALID, "bad dimensions");
    if (g.nq > 8) return fail(PYDSB_ERR_UNSUPPORTED, "at most 8 quadrature states");
    if (g.nce > 4) return fail(PYDSB_ERR_UNSUPPORTED, "at most 4 control functions");
    if (!resolve_table(secs, "RFQ_", g.nq, 8, g.fq_o, &g.fq_mask, US))
        return fail(PYDSB_ERR_INVALID, "bad fq reference table");
    {
        const double rtol = reinterpret_cast<const double*>(find(secs, "DBLS")->data)[0];
        g.rtol2 = rtol * rtol;
    }
    const Section* adot = find(secs, "ADOT");
    const Section* bw = find(secs, "BWTS");
    const Section* bigm = find(secs, "BIGM");
    if (adot->count != 16 || bw->count != 3 || (int)bigm->count < g.ng) return fail(PYDSB_ERR_INVALID, "bad collocation tables");
    memcpy(g.adot, adot->data, 16 * sizeof(double));
    memcpy(g.bw, bw->data, 3 * sizeof(double));
    for (int k = 0; k < g.ng; ++k) {
        const double M = reinterpret_cast<const double*>(bigm->data)[k];
        if (!(M > 0.0)) return fail(PYDSB_ERR_INVALID, "big-M constants must be positive");
        g.bigm_inv[k] = 1.0 / M;
    }
    const Section *tp = find(secs, "TPRO"), *te = find(secs, "TELM"), *tn = find(secs, "TPNT"), *tg = find(secs, "TGEE");
    g.n_pro = tp->count; g.n_elem = te->count; g.n_pt = tn->count; g.n_g = tg->count;
    std::vector<uint64_t> code;
    for (const Section* s : {tp, te, tg}) {
        const uint64_t* c = reinterpret_cast<const uint64_t*>(s->data);
        code.insert(code.end(), c, c + s->count);
    }
    for (uint64_t w : code)
        if ((w & 0xFF) >= OP_COUNT) return fail(PYDSB_ERR_INVALID, "unknown opcode in tape");
    // pre-decode the point tape: per operand the byte offsets of the three collocation points
    if (sens) {
        // program 1: the data ops of the four tapes for sens_kernel's interpreter
        AsmInput in{};
        in.pro = reinterpret_cast<const uint64_t*>(tp->data); in.n_pro = g.n_pro;
        in.elem = reinterpret_cast<const uint64_t*>(te->data); in.n_elem = g.n_elem;
        in.g = reinterpret_cast<const uint64_t*>(tg->data); in.n_g = g.n_g;
        in.pt = reinterpret_cast<const uint64_t*>(tn->data);
        in.n_units = g.n_units; in.spt = 1; in.img = &g;
        Assembler as;
        if (!as.assemble_segments(in, g.n_pt, g.seg)) return fail(PYDSB_ERR_UNSUPPORTED, "assembler: " + as.err);
        P.xprog.swap(as.words);
        g.n_prog = (int)(P.xprog.size() / 4);
    }
    g.zero_off = 0; g.zero_slot = 0;
    if (!sens) {
        // a 1-state block with an identically zero Jacobian reads it from a slot INIT keeps at 0.0
        bool need_zero = false;
        for (int b = 0; b < g.n_blocks; ++b) {
            if (P.poly[b].is_poly) need_zero = need_zero || !(P.poly[b].present[0] && P.poly[b].present[1] && P.poly[b].present[2]);
            else need_zero = need_zero || (g.blk_ns[b] == 1 && !(g.blk_jmask[b] & 1u));
        }
        if (need_zero) { g.zero_slot = g.n_units; g.zero_off = (unsigned)g.n_units * US; }
    }
    if (!sens) {
        // program 0: assemble the unified per-scenario program of feas_kernel
        AsmInput in{};
        in.pro = reinterpret_cast<const uint64_t*>(tp->data); in.n_pro = g.n_pro;
        in.elem = reinterpret_cast<const uint64_t*>(te->data); in.n_elem = g.n_elem;
        in.g = reinterpret_cast<const uint64_t*>(tg->data); in.n_g = g.n_g;
        in.pt = reinterpret_cast<const uint64_t*>(tn->data);
        in.n_blocks = g.n_blocks; in.blk_npt = g.blk_npt; in.blk_ns = g.blk_ns; in.blk_lin = g.blk_lin;
        in.blk_flops = g.blk_flops; in.q_npt = g.q_npt;
        int sum_pt = g.q_npt;
        for (int b = 0; b < g.n_blocks; ++b) sum_pt += g.blk_npt[b];
        if (sum_pt != g.n_pt) return fail(PYDSB_ERR_INVALID, "block tape lengths do not add up to the point tape");
        in.nce = g.nce; in.ctrl_slot = g.ctrl_slot; in.nq = g.nq; in.qe_slot = g.qe_slot; in.n_units = g.n_units;
        const Section* rfq = find(secs, "RFQ_");
        if ((int)rfq->count < g.nq) return fail(PYDSB_ERR_INVALID, "bad fq reference table");
        in.fq_refs = reinterpret_cast<const uint16_t*>(rfq->data);
        in.img = &g;
        in.nfe = g.nfe;
        in.cmap = reinterpret_cast<const uint16_t*>(find(secs, "CMAP")->data);
        if ((int)find(secs, "CMAP")->count < g.nfe * g.nce) return fail(PYDSB_ERR_INVALID, "control map shorter than nfe x nce");
        in.poly = P.poly;
        in.spt = g.spt;
        in.x0_refs = reinterpret_cast<const uint16_t*>(find(secs, "RX0_")->data);
        in.q0_refs = reinterpret_cast<const uint16_t*>(find(secs, "RQ0_")->data);
        in.par_refs = reinterpret_cast<const uint16_t*>(find(secs, "RPAR")->data);
        in.g_refs = reinterpret_cast<const uint16_t*>(find(secs, "RG__")->data);
        if ((int)find(secs, "RX0_")->count < g.ns || (int)find(secs, "RQ0_")->count < g.nq ||
            (int)find(secs, "RPAR")->count < g.d_dim + g.p_dim || (int)find(secs, "RG__")->count < g.ng)
            return fail(PYDSB_ERR_INVALID, "reference table shorter than its dimension");
        in.n_par = g.d_dim + g.p_dim; in.ng = g.ng; in.xe_slot = g.xe_slot;
        in.nz = g.nz;
        if (g.nz > 0) {
            const Section* rzs = find(secs, "RZ__");
            if (!rzs || (int)rzs->count < g.nz) return fail(PYDSB_ERR_INVALID, "reference table shorter than its dimension");
            in.z_refs = reinterpret_cast<const uint16_t*>(rzs->data);
        }
        // HEAD[15] bit 0 (DaeModel(fused_chain=False)) or PYDSB_NO_CHAIN=1: keep the element loop interpreted
        const char* no_chain = getenv("PYDSB_NO_CHAIN");
        in.allow_chain = rchain || (!(H[15] & 1u) && !(no_chain && atoi(no_chain) != 0));
        Assembler as;
        // PYDSB_NO_COMPACT=1: keep the lowering's unit numbers (A/B switch); the sensitivity chain addresses its extra rows
        // behind n_units and keeps them too
        const char* no_compact = getenv("PYDSB_NO_COMPACT");
        as.compact = !rchain && !(no_compact && atoi(no_compact) != 0);
        if (!as.assemble(in)) return fail(PYDSB_ERR_UNSUPPORTED, "assembler: " + as.err);
        P.xprog.swap(as.words);
        P.chain_nb = as.chain_nb;
        P.chain_units = as.chain_nb > 0 ? as.chain_units : 0;
        if (rchain && P.xprog.size() / 4 > 2 * (size_t)MAX_PROG_R) { P.present = false; P.chain_nb = 0; return PYDSB_OK; }
        if (rchain) {
            // the sensitivity chain needs the fused shape (and what recourse_chain.cuh supports of it): otherwise this
            // program is simply absent and sens_kernel (program 1) evaluates
            bool ok = as.chain_nb > 0 && g.nz >= 1 && g.nz <= RC_NZ && g.nq <= RC_NQ && g.nfe <= RC_MAXFE && g.nce <= 4 && g.ng <= 8;
            const ChainDesc& cdx = as.chain;
            for (unsigned b = 0; ok && b < cdx.nb; ++b) {
                const unsigned m2 = cdx.blk[b].c[2].meta;
                ok = ok && !(m2 & CT_HAS_CTRL);
                for (int k = 0; k < 2; ++k) {
                    const unsigned mk = cdx.blk[b].c[k].meta;
                    ok = ok && !((mk & CT_HAS_CTRL) && ((mk >> 8) & 3u) != 0u);      // a control enters as narrow x control
                }
            }
            for (unsigned q = 0; ok && q < cdx.nq; ++q) {
                const unsigned mq = cdx.quad[q].meta;
                ok = ok && !((mq & CT_HAS_CTRL) && ((mq >> 8) & 3u) != 0u);
            }
            const Section *rgx = find(secs, "RGX_"), *rgq = find(secs, "RGQ_"), *rz = find(secs, "RZ__"), *zf = find(secs, "ZFIX");
            ok = ok && rgx && rgq && (int)rgx->count >= g.ng * g.ns && (int)rgq->count >= (g.nq ? g.ng * g.nq : 0) &&
                 rz && zf && (int)rz->count >= g.nz && (int)zf->count >= g.nz;
            if (!ok) { P.present = false; P.chain_nb = 0; return PYDSB_OK; }
            RChainDesc& rd = P.rchain;
            rd = RChainDesc{};
            rd.ch = as.chain;
            chain_shape_bits(rd.ch, P.shape_sb, P.shape_sq);
            auto enc = [&](unsigned ref) -> unsigned {       // narrow reference -> slot offset / constant-pool offset / none
                if (ref == REF_NONE) return RC_REF_NONE;
                if (ref & REF_CONST) return RC_REF_CONST | (((ref & REF_MASK) + 3u) * 8u);
                return (ref & REF_MASK) * US;
            };
            const uint16_t* GX = reinterpret_cast<const uint16_t*>(rgx->data);
            const uint16_t* GQ = reinterpret_cast<const uint16_t*>(rgq->data);
            for (int k = 0; k < 8; ++k) {
                for (int b = 0; b < CHAIN_MAXB; ++b) rd.gx_off[k][b] = RC_REF_NONE;
                for (int q = 0; q < RC_NQ; ++q) rd.gq_off[k][q] = RC_REF_NONE;
            }
            for (int k = 0; k < g.ng; ++k) {
                for (unsigned b = 0; b < rd.ch.nb; ++b) {
                    const int st = (int)(rd.ch.blk[b].x_off / US) - g.x_slot;          // the block's state: x_slot + 3 st
                    rd.gx_off[k][b] = enc(GX[k * g.ns + st / 3]);
                }
                for (int q = 0; q < g.nq; ++q) rd.gq_off[k][q] = enc(GQ[k * g.nq + q]);
            }
            const uint16_t* cm = reinterpret_cast<const uint16_t*>(find(secs, "CMAP")->data);
            unsigned live = 0;
            for (int e = 0; e < g.nfe; ++e) {
                for (int k = 0; k < g.nce; ++k) { rd.cmap[e][k] = (unsigned char)cm[e * g.nce + k]; live |= 1u << cm[e * g.nce + k]; }
                rd.live[e] = live;
            }
            rd.nce = (unsigned)g.nce;
            const uint16_t* RZ = reinterpret_cast<const uint16_t*>(rz->data);
            const uint16_t* ZF = reinterpret_cast<const uint16_t*>(zf->data);
            for (int c = 0; c < g.nz; ++c) {
                rd.zfix_ref[c] = RC_REF_NONE;
                if (ZF[c]) { rd.zfix_mask |= 1u << c; rd.zfix_ref[c] = enc(RZ[c]); }
            }
        } else if (P.chain_nb > 0) {
            P.chain = as.chain;                              // what the kernel reads from the constant bank
            const ChainDesc& cd = P.chain;
            chain_shape_bits(cd, P.shape_sb, P.shape_sq);
            // the trajectory variant of the kernel stores the points element by element: it runs the interpreted loop
            in.allow_chain = false;
            Assembler as2;
            if (!as2.assemble(in)) return fail(PYDSB_ERR_UNSUPPORTED, "assembler: " + as2.err);
            P.xprog_interp.swap(as2.words);
        }
        g.n_prog = (int)(P.xprog.size() / 4);             // uint4 units
    }
    const Section* cn = find(secs, "CNST");
    g.n_const = cn->count;

    std::vector<uint16_t> tabs;
    auto push = [&](const char* tag, size_t n) -> bool {
        const Section* s = find(secs, tag);
        if (!s || s->count < n) return false;
        const uint16_t* v = reinterpret_cast<const uint16_t*>(s->data);
        tabs.insert(tabs.end(), v, v + n);
        return true;
    };
    bool ok = push("RPAR", (size_t)g.d_dim + g.p_dim) && push("RX0_", g.ns) && push("RQ0_", g.nq) && push("RZ__", g.nz) &&
              push("ZFIX", g.nz) && push("RG__", g.ng) && push("CMAP", (size_t)g.nfe * g.nce);
    if (!ok) return fail(PYDSB_ERR_INVALID, "reference table shorter than its dimension");
    if (sens) {
        if (g.nz > MAX_NZ) return fail(PYDSB_ERR_UNSUPPORTED, "at most 16 control values for the recourse optimiser");
        if (g.n_blocks != 1 || g.ns > 3) return fail(PYDSB_ERR_UNSUPPORTED, "the sensitivity program is one coupled block of <= 3 states");
        g.tri = 1;                                          // no Jacobian entry above the diagonal?
        for (int n = 0; n < g.ns; ++n)
            for (int m2 = n + 1; m2 < g.ns; ++m2)
                if ((g.blk_jmask[0] >> (n * g.ns + m2)) & 1u) g.tri = 0;
        {
            // compact lists of the non-zero sensitivity integrands: df/dz (ns x nce), dfq/dx (nq x ns), dfq/dz (nq x nce)
            unsigned tmp[32][3], mask = 0;
            int n_e = 0;
            auto add = [&](const char* tag, int rows, int cols, int& count) -> bool {
                if (rows * cols > 32 || !resolve_table(secs, tag, rows * cols, 32, tmp, &mask, US)) return false;
                count = 0;
                for (int i = 0; i < rows * cols; ++i) {
                    if (!((mask >> i) & 1u)) continue;
                    if (n_e >= MAX_SENS_ENTRIES) return false;
                    SensEntry& en = g.sens_e[n_e++];
                    en.row = (unsigned short)(i / cols); en.col = (unsigned short)(i % cols);
                    for (int p3 = 0; p3 < 3; ++p3) en.o[p3] = tmp[i][p3];
                    ++count;
                }
                return true;
            };
            if (!add("RFZ_", g.ns, g.nce, g.n_fz) || !add("RFQX", g.nq, g.ns, g.n_fqx) || !add("RFQZ", g.nq, g.nce, g.n_fqz))
                return fail(PYDSB_ERR_INVALID, "bad sensitivity reference table");
        }
        if (!push("RGX_", (size_t)g.ng * g.ns) || !push("RGQ_", (size_t)g.ng * g.nq))
            return fail(PYDSB_ERR_INVALID, "bad dg reference table");
        // live[e]: controls that have acted in an element <= e
        const uint16_t* cmap = reinterpret_cast<const uint16_t*>(find(secs, "CMAP")->data);
        unsigned live = 0;
        for (int e = 0; e < g.nfe; ++e) {
            for (int k = 0; k < g.nce; ++k) live |= 1u << cmap[e * g.nce + k];
            tabs.push_back((uint16_t)live);
        }
    }
    g.n_tabs = (int)tabs.size();

    // shared-memory carve-up
    int o = 128;                                            // mbarrier + reduction scratch
    g.off_const = o; o += (g.n_const + 3) * 8 * g.spt;      // chain programs: SPT-wide entries after three specials
    g.off_tabs = o; o = align_up(o + g.n_tabs * 2, 128);
    int land = g.spt * BLOCK * g.p_dim * 8;
    if (rchain) {
        land = 0;
        if (g.n_units < 1) return fail(PYDSB_ERR_UNSUPPORTED, "chain program too small for the reduction scratch");
    }
    if (sens) {
        land = 0;                                           // sens_kernel loads theta directly
        // the reduction scratch of the shared modes (rows of BLOCK doubles) overlays the program's slots
        if (g.n_units < 1) return fail(PYDSB_ERR_UNSUPPORTED, "sensitivity program too small for the reduction scratch");
    }
    g.off_land = o; o = align_up(o + land, 128);
    g.off_slots = o;
    int units = g.n_units + (g.zero_off ? 1 : 0);
    if (sens) units += recourse_sens_rows(g.ns, g.nq, g.ng) * g.nz;     // Sx, Sq (later Dc + one row) of sens_kernel
    if (rchain) units += (g.ng + 1) * g.nz;                             // Dc + one row (the sensitivities live in registers)
    P.chain_smem_bytes = (P.chain_nb > 0 && P.chain_units > 0) ? o + P.chain_units * g.spt * BLOCK * 8 : 0;
    o += units * g.spt * BLOCK * 8;
    g.smem_bytes = o;

    if (!upload) return PYDSB_OK;                           // host-only assembly (pydsb_assemble)
    cudaError_t e = cudaMalloc(&P.d_consts, (size_t)g.n_const * 8 + 8);
    if (e == cudaSuccess) e = cudaMemcpy(P.d_consts, cn->data, (size_t)g.n_const * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&P.d_tabs, tabs.size() * 2 + 8);
    if (e == cudaSuccess) e = cudaMemcpy(P.d_tabs, tabs.data(), tabs.size() * 2, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail(PYDSB_ERR_CUDA, std::string("model upload: ") + cudaGetErrorString(e));
    g.consts = static_cast<const double*>(P.d_consts);
    g.tabs = static_cast<const uint16_t*>(P.d_tabs);
    P.present = true;
    return PYDSB_OK;
}

// Split a blob into the section lists of its (one or two) programs.
int split_blob(const void* blob, size_t nbytes, Sections (&progs)[3], int& n_progs)
{
    if (!blob || nbytes < 16) return fail(PYDSB_ERR_INVALID, "null or truncated blob");
    const unsigned char* p = static_cast<const unsigned char*>(blob);
    uint32_t magic, ver, nsec;
    memcpy(&magic, p, 4); memcpy(&ver, p + 4, 4); memcpy(&nsec, p + 8, 4);
    if (magic != 0x42534450u || ver != 1u) return fail(PYDSB_ERR_INVALID, "bad blob magic/version");
    int cur = 0;
    size_t off = 16;
    static const size_t dsz[4] = {4, 8, 8, 2};
    for (uint32_t i = 0; i < nsec; ++i) {
        if (off + 16 > nbytes) return fail(PYDSB_ERR_INVALID, "truncated section header");
        Section s{};
        memcpy(s.tag, p + off, 4);
        memcpy(&s.dtype, p + off + 4, 4);
        memcpy(&s.count, p + off + 8, 4);
        if (s.dtype > 3) return fail(PYDSB_ERR_INVALID, "bad section dtype");
        off += 16;
        size_t nb = (size_t)s.count * dsz[s.dtype];
        nb = (nb + 7) / 8 * 8;
        if (off + nb > nbytes) return fail(PYDSB_ERR_INVALID, "truncated section payload");
        s.data = p + off;
        off += nb;
        if (strncmp(s.tag, "PSEP", 4) == 0) {             // payload: the number of the program that follows
            uint32_t idx = 0;
            if (s.count >= 1) memcpy(&idx, s.data, 4);
            if (idx < 1 || idx > 2) return fail(PYDSB_ERR_INVALID, "bad program separator in blob");
            cur = (int)idx;
            if (cur + 1 > n_progs) n_progs = cur + 1;
            continue;
        }
        progs[cur].push_back(s);
    }
    if (n_progs < 1) n_progs = 1;
    return PYDSB_OK;
}

int ensure_partial(Model* m, size_t n)
{
    if (n <= m->partial_cap) return 0;
    if (m->d_partial) cudaFree(m->d_partial);
    m->d_partial = nullptr;
    m->partial_cap = 0;
    CUDA_TRY(cudaMalloc(&m->d_partial, n * sizeof(double)));
    m->partial_cap = n;
    return 0;
}

}  // namespace

extern "C" {

const char* pydsb_last_error(void) { return g_err.c_str(); }
int pydsb_version(void) { return 1; }

int pydsb_model_create(const void* blob, size_t nbytes, int device, void** handle)
{
    if (!blob || !handle || nbytes < 16) return fail(PYDSB_ERR_INVALID, "null or truncated blob");
    Sections progs[3];
    int n_progs = 0;
    {
        const int prc = split_blob(blob, nbytes, progs, n_progs);
        if (prc != PYDSB_OK) return prc;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PYDSB_ERR_NO_DEVICE, "no CUDA device visible: pyds_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(PYDSB_ERR_INVALID, "device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    Model* m = new Model();
    m->device = device;
    int rc = build_program(progs[0], 0, m->prog[0]);
    if (rc == PYDSB_OK && !progs[1].empty()) rc = build_program(progs[1], 1, m->prog[1]);
    if (rc == PYDSB_OK && !progs[2].empty()) rc = build_program(progs[2], 2, m->prog[2]);
    if (rc != PYDSB_OK) {
        std::string msg = g_err;
        pydsb_model_destroy(m);
        return fail(rc, msg);
    }
    const ExecImage& g = m->prog[0].img;
    // control bounds
    {
        const Section *zl = find(progs[0], "ZLOW"), *zh = find(progs[0], "ZHIG"), *zf = find(progs[0], "ZFIX");
        const int nz = g.nz;
        m->z_lo.assign(nz, 0.0); m->z_hi.assign(nz, 0.0); m->z_fix.assign(nz > 0 ? nz : 1, 0);
        if (nz > 0) {
            if (!zl || !zh || (int)zl->count < nz || (int)zh->count < nz || (int)zf->count < nz) {
                pydsb_model_destroy(m);
                return fail(PYDSB_ERR_INVALID, "control bound tables shorter than nz");
            }
            memcpy(m->z_lo.data(), zl->data, nz * sizeof(double));
            memcpy(m->z_hi.data(), zh->data, nz * sizeof(double));
            for (int i = 0; i < nz; ++i) m->z_fix[i] = (unsigned char)reinterpret_cast<const uint16_t*>(zf->data)[i];
        }
    }
    cudaDeviceProp prop{};
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    const int nzb = g.nz > 0 ? g.nz : 1;
    if (e == cudaSuccess) e = cudaMalloc(&m->d_zlo, nzb * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&m->d_zhi, nzb * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&m->d_zfix, nzb);
    if (e == cudaSuccess && g.nz > 0) e = cudaMemcpy(m->d_zlo, m->z_lo.data(), g.nz * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && g.nz > 0) e = cudaMemcpy(m->d_zhi, m->z_hi.data(), g.nz * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_zfix, m->z_fix.data(), nzb, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_counters, PYDSB_CNT_COUNT * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(m->d_counters, 0, PYDSB_CNT_COUNT * sizeof(unsigned long long));
    if (e == cudaSuccess) {
        for (int q = 0; q < 3; ++q)
            if (m->prog[q].present && (size_t)m->prog[q].img.smem_bytes > prop.sharedMemPerBlockOptin) {
                pydsb_model_destroy(m);
                return fail(PYDSB_ERR_UNSUPPORTED, "model needs more shared memory per CTA than the device offers");
            }
    }
    FeasFn fn = feas_for(image_maxns(g), g.spt, &m->prog[0]);
    const int feas_smem = m->prog[0].chain_smem_bytes ? m->prog[0].chain_smem_bytes : g.smem_bytes;
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, feas_smem);
    int bps = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, BLOCK, feas_smem);
    cudaFuncAttributes fa{};
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, fn);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_async, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        std::string msg = std::string("model upload: ") + cudaGetErrorString(e);
        pydsb_model_destroy(m);
        return fail(PYDSB_ERR_CUDA, msg);
    }
    m->num_sms = prop.multiProcessorCount;
    m->blocks_per_sm = bps > 0 ? bps : 1;
    m->regs = fa.numRegs;
    *handle = m;
    return PYDSB_OK;
}

// Host-only: the bounding ellipsoid of the live points (pyds_b200/ns.py): mean, Cholesky factor of the covariance and the
// largest Mahalanobis radius of the points scaled to the unit box.  Two passes over the points.
int pydsb_ns_ellipsoid(const double* live, long long n, int dim, const double* lo, const double* hi, double* mu_out,
                       double* L_out, double* r_out)
{
    constexpr int MAXD = 32;
    if (!live || !lo || !hi || !mu_out || !L_out || !r_out || n < 1 || dim < 1 || dim > MAXD)
        return fail(PYDSB_ERR_INVALID, "bad ellipsoid arguments (1..32 design variables, at least one live point)");
    double inv_span[MAXD], mu[MAXD], C[MAXD][MAXD], L[MAXD][MAXD];
    for (int k = 0; k < dim; ++k) {
        if (!(hi[k] > lo[k])) return fail(PYDSB_ERR_INVALID, "empty design box");
        inv_span[k] = 1.0 / (hi[k] - lo[k]);
        mu[k] = 0.0;
    }
    for (long long i = 0; i < n; ++i)
        for (int k = 0; k < dim; ++k) mu[k] += (live[i * dim + k] - lo[k]) * inv_span[k];
    for (int k = 0; k < dim; ++k) mu[k] /= (double)n;
    for (int a = 0; a < dim; ++a)
        for (int b = 0; b < dim; ++b) C[a][b] = 0.0;
    if (n > dim) {
        for (long long i = 0; i < n; ++i) {
            double x[MAXD];
            for (int k = 0; k < dim; ++k) x[k] = (live[i * dim + k] - lo[k]) * inv_span[k] - mu[k];
            for (int a = 0; a < dim; ++a)
                for (int b = 0; b <= a; ++b) C[a][b] += x[a] * x[b];
        }
        for (int a = 0; a < dim; ++a)
            for (int b = 0; b <= a; ++b) C[a][b] = C[a][b] / (double)(n - 1) + (a == b ? 1e-12 : 0.0);
    } else {
        for (int a = 0; a < dim; ++a) C[a][a] = 0.25;       // too few points for a covariance: a fixed ball
    }
    for (int a = 0; a < dim; ++a)
        for (int b = 0; b <= a; ++b) {
            double s = C[a][b];
            for (int k = 0; k < b; ++k) s -= L[a][k] * L[b][k];
            if (a == b) {
                if (!(s > 0.0)) return fail(PYDSB_ERR_INVALID, "live-point covariance is not positive definite");
                L[a][a] = sqrt(s);
            } else {
                L[a][b] = s / L[b][b];
            }
        }
    double r2 = 0.0;                                         // max |L^-1 (u - mu)|^2 by forward substitution
    for (long long i = 0; i < n; ++i) {
        double y[MAXD], q = 0.0;
        for (int a = 0; a < dim; ++a) {
            double s = (live[i * dim + a] - lo[a]) * inv_span[a] - mu[a];
            for (int k = 0; k < a; ++k) s -= L[a][k] * y[k];
            y[a] = s / L[a][a];
            q += y[a] * y[a];
        }
        if (q > r2) r2 = q;
    }
    for (int a = 0; a < dim; ++a) {
        mu_out[a] = mu[a];
        for (int b = 0; b < dim; ++b) L_out[a * dim + b] = b <= a ? L[a][b] : 0.0;
    }
    *r_out = sqrt(r2);
    return PYDSB_OK;
}

// Host-only: the live-set update of one nested-sampling iteration (pyds_b200/ns.py), candidate by candidate.
int pydsb_ns_replace_worst(double* live_p, long long n_live, const double* cand_p, long long n_cand, double target,
                           long long* slot_out)
{
    if (n_live < 0 || n_cand < 0 || (n_live > 0 && !live_p) || (n_cand > 0 && (!cand_p || !slot_out)))
        return fail(PYDSB_ERR_INVALID, "bad live set / candidate arrays");
    if (n_live == 0) {
        for (long long c = 0; c < n_cand; ++c) slot_out[c] = -1;
        return PYDSB_OK;
    }
    // min-heap of (P, slot): the worst live point, the lowest slot among equal minima (what np.argmin would pick)
    typedef std::pair<double, long long> Entry;
    std::vector<Entry> heap((size_t)n_live);
    for (long long k = 0; k < n_live; ++k) heap[(size_t)k] = Entry(live_p[k], k);
    auto worse_last = [](const Entry& a, const Entry& b) { return a > b; };
    std::make_heap(heap.begin(), heap.end(), worse_last);
    for (long long c = 0; c < n_cand; ++c) {
        const double cp = cand_p[c];
        const Entry top = heap.front();
        slot_out[c] = -1;
        if (cp > top.first || (cp == top.first && cp >= target)) {
            std::pop_heap(heap.begin(), heap.end(), worse_last);
            heap.back() = Entry(cp, top.second);
            std::push_heap(heap.begin(), heap.end(), worse_last);
            live_p[top.second] = cp;
            slot_out[c] = top.second;
        }
    }
    return PYDSB_OK;
}

int pydsb_assemble(const void* blob, size_t nbytes, uint32_t* words, size_t cap, size_t* n_words, long long* info, int n_info)
{
    Sections progs[3];
    int n_progs = 0;
    int rc = split_blob(blob, nbytes, progs, n_progs);
    if (rc != PYDSB_OK) return rc;
    Program P;
    rc = build_program(progs[0], 0, P, /*upload=*/false);
    if (rc != PYDSB_OK) return rc;
    if (n_words) *n_words = P.xprog.size();
    if (words)
        for (size_t i = 0; i < P.xprog.size() && i < cap; ++i) words[i] = P.xprog[i];
    const ExecImage& g = P.img;
    const long long v[8] = {g.spt, P.chain_units ? P.chain_units : g.n_units + (g.zero_off ? 1 : 0),
                            P.chain_smem_bytes ? P.chain_smem_bytes : g.smem_bytes, g.n_prog, P.chain_nb,
                            (long long)P.shape_sb, (long long)P.shape_sq, chain_kernel_for(P, g.spt) ? 1 : 0};
    for (int i = 0; info && i < n_info && i < 8; ++i) info[i] = v[i];
    // info[8 ...]: the chain descriptor, word by word (ChainDesc, poly_chain.cuh) -- inspection and the CPU tests
    const uint32_t* cdw = reinterpret_cast<const uint32_t*>(&P.chain);
    for (int i = 8; info && i < n_info && i < 8 + (int)(sizeof(ChainDesc) / 4); ++i) info[i] = P.chain_nb > 0 ? cdw[i - 8] : 0;
    return PYDSB_OK;
}

int pydsb_model_destroy(void* handle)
{
    Model* m = static_cast<Model*>(handle);
    if (!m) return PYDSB_OK;
    cudaSetDevice(m->device);
    free_program(m->prog[0]);
    free_program(m->prog[1]);
    free_program(m->prog[2]);
    if (m->d_zlo) cudaFree(m->d_zlo);
    if (m->d_zhi) cudaFree(m->d_zhi);
    if (m->d_zfix) cudaFree(m->d_zfix);
    if (m->d_counters) cudaFree(m->d_counters);
    if (m->d_partial) cudaFree(m->d_partial);
    if (m->d_stage) cudaFree(m->d_stage);
    for (void* p : m->rws) if (p) cudaFree(p);
    if (m->h_pin) cudaFreeHost(m->h_pin);
    prof_collect(m);
    if (m->ev_async) cudaEventDestroy(m->ev_async);
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    delete m;
    return PYDSB_OK;
}

int pydsb_model_info(void* handle, long long* out, int n)
{
    Model* m = static_cast<Model*>(handle);
    if (!m || !out) return fail(PYDSB_ERR_INVALID, "null handle/out");
    const ExecImage& g = m->prog[0].img;
    long long v[PYDSB_INFO_COUNT] = {g.ns, g.nq, g.ng, g.d_dim, g.p_dim, g.nz, g.nfe, 3, g.n_units, g.smem_bytes,
                                     m->blocks_per_sm, m->num_sms, m->regs, g.spt, g.n_prog, m->prog[0].chain_nb,
                                     m->prog[2].present ? m->prog[2].chain_nb : 0};
    for (int i = 0; i < n && i < PYDSB_INFO_COUNT; ++i) out[i] = v[i];
    return PYDSB_OK;
}

}  // extern "C"

namespace {

// pydsb_eval, or with `traj` the trajectory variant (g_out = x_out, ind_out = failed_out)
int eval_impl(void* handle, const double* d, long long n_d, const double* theta, const double* w, long long n_t,
              const double* z, int z_mode, double* g_out, double* ind_out, double* P_out, bool traj, void* stream,
              double feas_tol = 0.0)
{
    Model* m = static_cast<Model*>(handle);
    if (!m) return fail(PYDSB_ERR_INVALID, "null handle");
    if (n_d < 0 || n_t < 0) return fail(PYDSB_ERR_INVALID, "negative size");
    if (n_d == 0) return PYDSB_OK;
    const ExecImage& img = m->prog[0].img;
    if (!theta || (!d && img.d_dim > 0)) return fail(PYDSB_ERR_INVALID, "null d/theta");
    if (z_mode < 0 || z_mode > 3) return fail(PYDSB_ERR_INVALID, "bad z_mode");
    if (z_mode != 0 && !z && img.nz > 0) return fail(PYDSB_ERR_INVALID, "z_mode needs a z array");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(m->device));
    if (n_t == 0) {                     // empty theta set: P = 0 for every design point
        if (P_out) CUDA_TRY(cudaMemsetAsync(P_out, 0, (size_t)n_d * sizeof(double), st));
        return PYDSB_OK;
    }
    if (st == m->own_stream && m->async_pending) {
        // an asynchronous call on the caller's stream may still be using the partial sums and the program window
        CUDA_TRY(cudaStreamWaitEvent(st, m->ev_async, 0));
        m->async_pending = false;
    }
    const long long tile = (long long)img.spt * BLOCK;               // theta rows per tile
    // a theta set of at most half a tile (the shipped scripts use 50 samples): tile / n_t design points share a tile
    const bool pack = !traj && 2 * n_t <= tile;
    const long long pack_k = pack ? tile / n_t : 1;
    const long long n_tiles_ll = pack ? 1 : (n_t + tile - 1) / tile;
    if (n_tiles_ll > 0x7fffffffLL / (BLOCK / 32)) return fail(PYDSB_ERR_INVALID, "n_t too large");
    const int n_tiles = (int)n_tiles_ll;                              // partial sums per design point
    const long long total = pack ? (n_d + pack_k - 1) / pack_k : n_d * n_tiles_ll;   // tiles = CTAs' work items
    const int n_part = pack ? 1 : n_tiles * (BLOCK / 32);              // partial sums per design point: one per warp and tile
    int rc = ensure_partial(m, (size_t)n_d * (size_t)n_part);
    if (rc) return rc;

    LaunchArgs la{};
    la.d = d; la.theta = theta; la.w = w; la.z = z; la.z_mode = z_mode;
    la.n_d = n_d; la.n_t = n_t; la.n_tiles = n_tiles;
    la.theta_tma_ok = (reinterpret_cast<uintptr_t>(theta) % 16 == 0) ? 1 : 0;
    la.partial = m->d_partial; la.g_out = g_out; la.ind_out = ind_out; la.counters = m->d_counters;
    if (!(feas_tol >= 0.0)) return fail(PYDSB_ERR_INVALID, "feasibility tolerance must be >= 0");
    for (int k = 0; k < img.ng; ++k) la.gtol[k] = feas_tol > 0.0 ? feas_tol / img.bigm_inv[k] : 0.0;

    const long long max_ctas = (long long)m->num_sms * m->blocks_per_sm;
    const int grid = (int)(total < max_ctas ? total : max_ctas);
    const int chain = m->prog[0].chain_nb;
    FeasFn fn = traj ? traj_for(image_maxns(img), img.spt)
                     : (pack ? pack_for(image_maxns(img), img.spt, &m->prog[0]) : feas_for(image_maxns(img), img.spt, &m->prog[0]));
    const std::vector<uint32_t>& xp = (traj && chain > 0) ? m->prog[0].xprog_interp : m->prog[0].xprog;
    // several models may share one template instantiation: (re)assert this model's smem size and
    // (re)load its program into the constant bank (stream ordered)
    const int smem = (!traj && chain > 0 && m->prog[0].chain_smem_bytes) ? m->prog[0].chain_smem_bytes : img.smem_bytes;
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_prog, xp.data(), xp.size() * sizeof(uint32_t), 0, cudaMemcpyHostToDevice, st));
    if (chain > 0 && !traj)
        CUDA_TRY(cudaMemcpyToSymbolAsync(c_chain, &m->prog[0].chain, sizeof(ChainDesc), 0, cudaMemcpyHostToDevice, st));
    {
        ProfScope ps(m, 0, st);
        fn<<<grid, BLOCK, smem, st>>>(img, la);
    }
    CUDA_TRY(cudaGetLastError());
    m->launches += 1;
    if (P_out) {
        const int tb = 128;
        reduce_partials_kernel<<<(unsigned)((n_d + tb - 1) / tb), tb, 0, st>>>(m->d_partial, n_part, n_d, P_out);
        CUDA_TRY(cudaGetLastError());
        m->launches += 1;
    }
    if (st != m->own_stream) {
        CUDA_TRY(cudaEventRecord(m->ev_async, st));
        m->async_pending = true;
    }
    return PYDSB_OK;
}

}  // namespace

extern "C" {

int pydsb_eval(void* handle, const double* d, long long n_d, const double* theta, const double* w, long long n_t,
               const double* z, int z_mode, double* g_out, double* ind_out, double* P_out, void* stream)
{
    return eval_impl(handle, d, n_d, theta, w, n_t, z, z_mode, g_out, ind_out, P_out, false, stream);
}

int pydsb_eval_tol(void* handle, const double* d, long long n_d, const double* theta, const double* w, long long n_t,
                   const double* z, int z_mode, double feas_tol, double* g_out, double* ind_out, double* P_out, void* stream)
{
    return eval_impl(handle, d, n_d, theta, w, n_t, z, z_mode, g_out, ind_out, P_out, false, stream, feas_tol);
}

int pydsb_state_solve(void* handle, const double* d, long long n_d, const double* theta, long long n_t, const double* z,
                      int z_mode, double* x_out, double* failed_out, void* stream)
{
    if (!x_out) return fail(PYDSB_ERR_INVALID, "null x_out");
    return eval_impl(handle, d, n_d, theta, nullptr, n_t, z, z_mode, x_out, failed_out, nullptr, true, stream);
}

int pydsb_eval_g(void* handle, const double* d, long long n_d, const double* theta, long long n_t, double* g_out,
                 void* stream)
{
    return pydsb_eval(handle, d, n_d, theta, nullptr, n_t, nullptr, PYDSB_Z_MODEL, g_out, nullptr, nullptr, stream);
}

int pydsb_efp(void* handle, const double* d, long long n_d, const double* theta, const double* w, long long n_t,
              double* P_out, void* stream)
{
    return pydsb_eval(handle, d, n_d, theta, w, n_t, nullptr, PYDSB_Z_MODEL, nullptr, nullptr, P_out, stream);
}

int pydsb_eval_host(void* handle, const double* d, long long n_d, const double* theta, const double* w, long long n_t,
                    const double* z, int z_mode, double* g_out, double* ind_out, double* P_out)
{
    Model* m = static_cast<Model*>(handle);
    if (!m) return fail(PYDSB_ERR_INVALID, "null handle");
    if (n_d < 0 || n_t < 0) return fail(PYDSB_ERR_INVALID, "negative size");
    if (n_d == 0) return PYDSB_OK;
    if (z_mode < 0 || z_mode > 3) return fail(PYDSB_ERR_INVALID, "bad z_mode");
    CUDA_TRY(cudaSetDevice(m->device));
    const ExecImage& g = m->prog[0].img;
    const size_t nd = (size_t)n_d * g.d_dim, nth = (size_t)n_t * g.p_dim, nw = w ? (size_t)n_t : 0;
    size_t nz = 0;
    if (z && z_mode == 1) nz = g.nz;
    if (z && z_mode == 2) nz = (size_t)n_d * g.nz;
    if (z && z_mode == 3) nz = (size_t)n_d * n_t * g.nz;
    const size_t ng = g_out ? (size_t)n_d * n_t * g.ng : 0, ni = ind_out ? (size_t)n_d * n_t : 0, np = P_out ? (size_t)n_d : 0;
    const size_t n_out = ng + ni + np;
    // layout (doubles): inputs padded so that theta starts 16-byte aligned
    const size_t o_d = 0, o_th = (nd + 1) / 2 * 2, o_w = o_th + nth, o_z = o_w + nw, o_out = (o_z + nz + 1) / 2 * 2;
    const size_t need = o_out + n_out;
    if (need > m->h_pin_cap) {
        if (m->h_pin) cudaFreeHost(m->h_pin);
        if (m->d_stage) cudaFree(m->d_stage);
        m->h_pin = nullptr; m->d_stage = nullptr; m->h_pin_cap = 0;
        CUDA_TRY(cudaMallocHost(&m->h_pin, need * sizeof(double)));
        CUDA_TRY(cudaMalloc(&m->d_stage, need * sizeof(double)));
        m->h_pin_cap = need;
    }
    double* hp = m->h_pin;
    double* dp = m->d_stage;
    if (nd) memcpy(hp + o_d, d, nd * sizeof(double));
    if (nth) memcpy(hp + o_th, theta, nth * sizeof(double));
    if (nw) memcpy(hp + o_w, w, nw * sizeof(double));
    if (nz) memcpy(hp + o_z, z, nz * sizeof(double));
    cudaStream_t st = m->own_stream;
    CUDA_TRY(cudaMemcpyAsync(dp, hp, o_out * sizeof(double), cudaMemcpyHostToDevice, st));
    double* dg = ng ? dp + o_out : nullptr;
    double* di = ni ? dp + o_out + ng : nullptr;
    double* dP = np ? dp + o_out + ng + ni : nullptr;
    int rc = pydsb_eval(handle, dp + o_d, n_d, dp + o_th, nw ? dp + o_w : nullptr, n_t, nz ? dp + o_z : nullptr,
                        nz ? z_mode : 0, dg, di, dP, st);
    if (rc) return rc;
    if (n_out) CUDA_TRY(cudaMemcpyAsync(hp + o_out, dp + o_out, n_out * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (ng) memcpy(g_out, hp + o_out, ng * sizeof(double));
    if (ni) memcpy(ind_out, hp + o_out + ng, ni * sizeof(double));
    if (np) memcpy(P_out, hp + o_out + ng + ni, np * sizeof(double));
    return PYDSB_OK;
}

}  // extern "C"

namespace {

// Shared body of pydsb_recourse_solve and pydsb_recourse_solve_global_dist.  `reduce` != NULL: the design points of the
// ONE global group are spread over several processes -- n_d rows here, n_d_total in all -- and after every evaluation
// the (stages x NV) statistics are summed over the processes through the callback before the (identical) update.
int recourse_solve_impl(void* handle, int per_scenario, const double* d, long long n_d, long long n_d_total,
                        const double* theta, const double* w, long long n_t, const double* z0, long long z0_rows,
                        const double* opts, double* z_out, double* phi_out, int* status_out, int* iters_out,
                        double* reduce_buf, pydsb_allreduce_fn reduce, void* user, void* stream)
{
    Model* m = static_cast<Model*>(handle);
    if (!m) return fail(PYDSB_ERR_INVALID, "null handle");
    if (!m->prog[1].present) return fail(PYDSB_ERR_UNSUPPORTED, "blob carries no sensitivity program");
    // evaluator: the chain with its sensitivities in registers (program 2, rsens_kernel) when the model has that shape,
    // else the coupled-block interpreter (program 1, sens_kernel); PYDSB_NO_RCHAIN=1 forces the latter
    const char* no_rc = getenv("PYDSB_NO_RCHAIN");
    const bool use_rc = m->prog[2].present && m->prog[2].chain_nb > 0 && !(no_rc && atoi(no_rc) != 0);
    const ExecImage& img = use_rc ? m->prog[2].img : m->prog[1].img;
    const int nz = img.nz;
    if (nz < 1) return fail(PYDSB_ERR_INVALID, "model has no control to optimise");
    if (n_d < 0 || n_t < 1) return fail(PYDSB_ERR_INVALID, "bad sizes");
    if (n_d == 0 && !reduce) return PYDSB_OK;
    if (!d && img.d_dim > 0 && n_d > 0) return fail(PYDSB_ERR_INVALID, "null d");
    if (!theta || !z_out) return fail(PYDSB_ERR_INVALID, "null theta/z_out");
    if (per_scenario < 0 || per_scenario > 2) return fail(PYDSB_ERR_INVALID, "mode must be 0, 1 or 2");
    if (reduce && (per_scenario != 2 || !reduce_buf || n_d_total < n_d || n_d_total < 1))
        return fail(PYDSB_ERR_INVALID, "the distributed solve is the global-control mode with a reduction buffer");
    const long long G = per_scenario == 1 ? n_d * n_t : (per_scenario == 2 ? 1 : n_d);
    if (z0 && !(z0_rows == 1 || z0_rows == G)) return fail(PYDSB_ERR_INVALID, "z0 must have 1 or G rows");
    // options: mu0, mu_min, lam0, step_tol, max_iter, check_every
    double o[6] = {1e-4, 1e-9, 1e-3, 1e-8, 200, 4};
    if (opts)
        for (int i = 0; i < 6; ++i)
            if (opts[i] > 0) o[i] = opts[i];
    const int max_iter = (int)o[4];
    int check_every = (int)o[5] < 1 ? 1 : (int)o[5];
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(m->device));

    const int n_tiles = (int)((n_t + BLOCK - 1) / BLOCK);
    const int NV = recourse_nv(nz);
    // tiles of the one global group run over the flattened (d, theta) index
    const long long global_tiles = (n_d * n_t + BLOCK - 1) / BLOCK;
    const long long acc_per_group_ll = per_scenario == 1 ? 1 : (per_scenario == 2 ? (global_tiles > 0 ? global_tiles : 1) : n_tiles);
    if (acc_per_group_ll > 0x7fffffffLL) return fail(PYDSB_ERR_INVALID, "grid too large");
    const int acc_per_group = (int)acc_per_group_ll;
    // workspace: doubles then ints
    const int n_stages = RECOURSE_STAGES;
    const size_t acc_dbl = per_scenario == 1 ? (size_t)G * recourse_nvp(img.ng, nz) : (size_t)G * acc_per_group * n_stages * NV;
    const size_t red_dbl = per_scenario == 2 ? (size_t)n_stages * NV : 0;   // the global group's summed statistics
    const size_t nd_dbl = (size_t)G * (2 * nz + 2 + nz + (size_t)nz * nz + 2) + acc_dbl + red_dbl;
    const bool compact = per_scenario == 1;      // running scenarios are compacted after every iteration
    const size_t nd_int = (size_t)G * 3 + 2;
    double* ws = nullptr;
    int* wsi = nullptr;
    long long* act = nullptr;                    // two active lists (per-scenario mode)
    const size_t want[3] = {nd_dbl * sizeof(double), nd_int * sizeof(int), compact ? 2 * (size_t)G * sizeof(long long) : 0};
    // Workspaces are kept in the model between calls (the nested-sampling loop calls with a few hundred groups thousands
    // of times; a per-scenario batch of 1e7 groups needs 9 GB, and allocating / releasing that much per call costs
    // hundreds of milliseconds on a busy allocator -- 0.05 s of kernels measured at 0.49 s).  Up to 16 GiB each; released
    // by pydsb_model_destroy.
    constexpr size_t KEEP = (size_t)16 << 30;
    void* got[3] = {nullptr, nullptr, nullptr};
    bool owned[3] = {false, false, false};       // allocated for this call only
    for (int i = 0; i < 3; ++i) {
        if (want[i] == 0) continue;
        if (want[i] <= m->rws_cap[i]) { got[i] = m->rws[i]; continue; }
        void* p = nullptr;
        if (cudaMalloc(&p, want[i]) != cudaSuccess) {
            cudaGetLastError();
            for (int k = 0; k < i; ++k) if (owned[k]) cudaFree(got[k]);
            return fail(PYDSB_ERR_CUDA, "cudaMalloc(recourse workspace)");
        }
        if (want[i] <= KEEP) {
            if (m->rws[i]) cudaFree(m->rws[i]);
            m->rws[i] = p; m->rws_cap[i] = want[i];
        } else {
            owned[i] = true;
        }
        got[i] = p;
    }
    ws = static_cast<double*>(got[0]); wsi = static_cast<int*>(got[1]); act = static_cast<long long*>(got[2]);
    double* z_cur = ws;
    double* z_try = z_cur + (size_t)G * nz;
    double* F_cur = z_try + (size_t)G * nz;
    double* raw_cur = F_cur + G;
    double* gr = raw_cur + G;
    double* Hh = gr + (size_t)G * nz;
    double* lam = Hh + (size_t)G * nz * nz;
    double* mu = lam + G;
    double* acc = mu + G;
    double* red = reduce_buf ? reduce_buf : acc + acc_dbl;
    int* status = wsi;
    int* first = status + G;
    int* iters = first + G;
    int* counts = iters + G;                     // [2]: groups still running after the even / odd iterations

    auto cleanup = [&]() { for (int i = 0; i < 3; ++i) if (owned[i]) cudaFree(got[i]); };
#define CUDA_TRY_WS(expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            cleanup();                                                                            \
            return fail(PYDSB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
        }                                                                                         \
    } while (0)

    const int tb = 128;
    const unsigned gb = (unsigned)((G + tb - 1) / tb);
    lm_init_kernel<<<gb, tb, 0, st>>>(G, nz, z0, (int)(z0 ? z0_rows : 0), z_cur, z_try, F_cur, raw_cur, lam, mu, status,
                                      first, iters, o[2], o[0]);
    CUDA_TRY_WS(cudaGetLastError());
    SensFn fn = use_rc ? rsens_for(m->prog[2]) : sens_for(img.ns, img.tri != 0);
    CUDA_TRY_WS(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, img.smem_bytes));
    if (use_rc) {
        CUDA_TRY_WS(cudaMemcpyToSymbolAsync(c_prog_r, m->prog[2].xprog.data(), m->prog[2].xprog.size() * sizeof(uint32_t), 0,
                                            cudaMemcpyHostToDevice, st));
        CUDA_TRY_WS(cudaMemcpyToSymbolAsync(c_rchain, &m->prog[2].rchain, sizeof(RChainDesc), 0, cudaMemcpyHostToDevice, st));
    } else {
        CUDA_TRY_WS(cudaMemcpyToSymbolAsync(c_prog_s, m->prog[1].xprog.data(), m->prog[1].xprog.size() * sizeof(uint32_t), 0,
                                            cudaMemcpyHostToDevice, st));
    }
    RecourseArgs ra{};
    ra.d = d; ra.theta = theta; ra.w = w; ra.n_d = n_d; ra.n_d_total = n_d_total; ra.n_t = n_t; ra.n_tiles = n_tiles;
    ra.per_scenario = per_scenario;
    ra.z_try = z_try; ra.status = status; ra.mu = mu; ra.acc = acc; ra.counters = m->d_counters;
    ra.sens_slot = img.n_units;
    ra.mu_min = o[1]; ra.n_stages = n_stages;
    LmArgs la{};
    la.n_groups = G; la.nz = nz; la.ng = img.ng; la.n_tiles = acc_per_group; la.per_scenario = per_scenario == 1 ? 1 : 0;
    la.n_stages = n_stages; la.acc = acc; la.lo = m->d_zlo; la.hi = m->d_zhi;
    la.zfix = m->d_zfix; la.z_cur = z_cur; la.z_try = z_try; la.F_cur = F_cur; la.raw_cur = raw_cur; la.gr = gr; la.H = Hh;
    la.lam = lam; la.mu = mu; la.status = status; la.first = first; la.iters = iters;
    la.mu_min = o[1]; la.lam0 = o[2]; la.step_tol = o[3]; la.max_iter = max_iter;
    const long long total_tiles = per_scenario == 2 ? global_tiles : n_d * (long long)n_tiles;
    // small grids are bound by the latency of one launch, not by throughput: look at the running count every iteration
    if (!(opts && opts[5] > 0) && total_tiles <= 4096) check_every = 1;
    if (reduce) check_every = 1;                 // every process must take the same decisions at the same iterations
    if (total_tiles > 0x7fffffffLL || G > 0x7fffffffLL) { cleanup(); return fail(PYDSB_ERR_INVALID, "grid too large"); }
    const char* no_fused = getenv("PYDSB_NO_FUSED_LM");
    if (use_rc && per_scenario == 0 && n_tiles == 1 && nz <= RC_NZ && !reduce && !(no_fused && atoi(no_fused) != 0)) {
        // a shared-control group whose theta set fits one tile (the shipped scripts: 50 samples): the whole optimisation of
        // the group runs inside ONE launch, a CTA per group (rfused_kernel) -- no launch pair, no count read-back per iteration
        RFusedFn ff = rfused_for(m->prog[2]);
        CUDA_TRY_WS(cudaFuncSetAttribute(ff, cudaFuncAttributeMaxDynamicSharedMemorySize, img.smem_bytes));
        ra.active = nullptr; ra.n_active = nullptr;
        la.active_in = nullptr; la.n_active_in = nullptr; la.active_out = nullptr; la.n_running = counts;
        ra.n_stages = 1; la.n_stages = 1;            // one smoothing stage per emission, the next one on request
        {
            ProfScope ps(m, 1, st);
            ff<<<(unsigned)G, BLOCK, img.smem_bytes, st>>>(img, ra, la);
        }
        CUDA_TRY_WS(cudaGetLastError());
        m->launches += 1;
    } else {
    // every group stops after max_iter evaluations (status 3) at the latest
    const long long* act_in = nullptr;           // null: every group (first iteration; always in the shared modes)
    const int* n_in = nullptr;
    long long upper = G;                         // host-known upper bound of the groups still running
    for (int it = 0; it < max_iter + 1; ++it) {
        int* n_out = counts + (it & 1);
        CUDA_TRY_WS(cudaMemsetAsync(n_out, 0, sizeof(int), st));
        ra.active = act_in; ra.n_active = n_in;
        const unsigned sens_grid = compact ? (unsigned)((upper + BLOCK - 1) / BLOCK) : (unsigned)total_tiles;
        if (sens_grid > 0) {
            {
                ProfScope ps(m, 1, st);
                fn<<<sens_grid, BLOCK, img.smem_bytes, st>>>(img, ra);
            }
            CUDA_TRY_WS(cudaGetLastError());
        }
        if (per_scenario == 2) {
            // the tile partials of the one global group -> one (stages x NV) vector; with several processes the
            // caller then sums it over the processes
            const int snv = n_stages * NV;
            sum_tiles_kernel<<<snv, 128, 0, st>>>(acc, (int)total_tiles, snv, red);
            CUDA_TRY_WS(cudaGetLastError());
            if (reduce) {
                CUDA_TRY_WS(cudaStreamSynchronize(st));
                if (reduce(red, snv, user) != 0) { cleanup(); return fail(PYDSB_ERR_INVALID, "all-reduce callback failed"); }
            }
            la.acc = red; la.n_tiles = 1;
            m->launches += 1;
        }
        la.active_in = act_in; la.n_active_in = n_in; la.n_running = n_out;
        la.active_out = compact ? act + (size_t)(it & 1) * G : nullptr;
        {
            ProfScope ps(m, 2, st);
            if (nz <= 8 && la.per_scenario && img.ng <= 2) lm_update_kernel<8, 2><<<(unsigned)((upper + tb - 1) / tb), tb, 0, st>>>(la);
            else if (nz <= 8) lm_update_kernel<8><<<(unsigned)((upper + tb - 1) / tb), tb, 0, st>>>(la);
            else lm_update_kernel<0><<<(unsigned)((upper + tb - 1) / tb), tb, 0, st>>>(la);
        }
        CUDA_TRY_WS(cudaGetLastError());
        m->launches += 2;
        if (compact) { act_in = la.active_out; n_in = n_out; }
        if ((it + 1) % check_every == 0 || it == max_iter) {
            int running = 0;
            CUDA_TRY_WS(cudaMemcpyAsync(&running, n_out, sizeof(int), cudaMemcpyDeviceToHost, st));
            CUDA_TRY_WS(cudaStreamSynchronize(st));
            if (running == 0) break;
            if (compact) upper = running;
        }
    }
    }
    CUDA_TRY_WS(cudaMemcpyAsync(z_out, z_cur, (size_t)G * nz * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (phi_out) CUDA_TRY_WS(cudaMemcpyAsync(phi_out, raw_cur, (size_t)G * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (status_out) CUDA_TRY_WS(cudaMemcpyAsync(status_out, status, (size_t)G * sizeof(int), cudaMemcpyDeviceToDevice, st));
    if (iters_out) CUDA_TRY_WS(cudaMemcpyAsync(iters_out, iters, (size_t)G * sizeof(int), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY_WS(cudaStreamSynchronize(st));
    cleanup();
#undef CUDA_TRY_WS
    return PYDSB_OK;
}

}  // namespace

extern "C" {

int pydsb_recourse_solve(void* handle, int per_scenario, const double* d, long long n_d, const double* theta,
                         const double* w, long long n_t, const double* z0, long long z0_rows, const double* opts,
                         double* z_out, double* phi_out, int* status_out, int* iters_out, void* stream)
{
    return recourse_solve_impl(handle, per_scenario, d, n_d, n_d, theta, w, n_t, z0, z0_rows, opts, z_out, phi_out, status_out,
                               iters_out, nullptr, nullptr, nullptr, stream);
}

int pydsb_recourse_solve_global_dist(void* handle, const double* d, long long n_d_local, long long n_d_total,
                                     const double* theta, const double* w, long long n_t, const double* z0,
                                     const double* opts, double* z_out, double* phi_out, int* status_out, int* iters_out,
                                     double* reduce_buf, pydsb_allreduce_fn reduce, void* user, void* stream)
{
    if (!reduce) return fail(PYDSB_ERR_INVALID, "null all-reduce callback");
    return recourse_solve_impl(handle, 2, d, n_d_local, n_d_total, theta, w, n_t, z0, z0 ? 1 : 0, opts, z_out, phi_out,
                               status_out, iters_out, reduce_buf, reduce, user, stream);
}

int pydsb_get_counters(void* handle, unsigned long long* out, int n, int reset)
{
    Model* m = static_cast<Model*>(handle);
    if (!m || !out) return fail(PYDSB_ERR_INVALID, "null handle/out");
    CUDA_TRY(cudaSetDevice(m->device));
    CUDA_TRY(cudaDeviceSynchronize());
    unsigned long long v[PYDSB_CNT_COUNT] = {0};
    CUDA_TRY(cudaMemcpy(v, m->d_counters, sizeof(v), cudaMemcpyDeviceToHost));
    v[PYDSB_CNT_LAUNCHES] = m->launches;
    for (int i = 0; i < n && i < PYDSB_CNT_COUNT; ++i) out[i] = v[i];
    if (reset) {
        CUDA_TRY(cudaMemset(m->d_counters, 0, sizeof(v)));
        m->launches = 0;
    }
    return PYDSB_OK;
}

int pydsb_set_profiling(void* handle, int on)
{
    Model* m = static_cast<Model*>(handle);
    if (!m) return fail(PYDSB_ERR_INVALID, "null handle");
    m->prof_on = on != 0;
    return PYDSB_OK;
}

int pydsb_get_profile(void* handle, double* out, int n, int reset)
{
    Model* m = static_cast<Model*>(handle);
    if (!m || !out) return fail(PYDSB_ERR_INVALID, "null handle/out");
    CUDA_TRY(cudaSetDevice(m->device));
    CUDA_TRY(cudaDeviceSynchronize());
    prof_collect(m);
    const double v[PYDSB_PROF_COUNT] = {m->prof_ms[0], (double)m->prof_n[0], m->prof_ms[1], (double)m->prof_n[1],
                                        m->prof_ms[2], (double)m->prof_n[2]};
    for (int i = 0; i < n && i < PYDSB_PROF_COUNT; ++i) out[i] = v[i];
    if (reset)
        for (int k = 0; k < 3; ++k) { m->prof_ms[k] = 0.0; m->prof_n[k] = 0; }
    return PYDSB_OK;
}

int pydsb_dfma_peak(int device, int repeats, double* tflops, double* ms_out)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PYDSB_ERR_NO_DEVICE, "no CUDA device visible");
    }
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop{};
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 4096;
    double* d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_out, (size_t)threads * blocks * sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    dfma_peak_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);       // warm-up
    CUDA_TRY(cudaDeviceSynchronize());
    float best = 1e30f;
    if (repeats < 1) repeats = 1;
    for (int r = 0; r < repeats; ++r) {
        CUDA_TRY(cudaEventRecord(e0));
        dfma_peak_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);
        CUDA_TRY(cudaEventRecord(e1));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double flops = 2.0 * 64.0 * (double)iters * threads * blocks;
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return PYDSB_OK;
}

}  // extern "C"
