// C-ABI of libpyds_b200.so (see include/pydsb.h).  Host side only: blob parsing, device image,
// launch configuration, pinned staging for the *_host entry points.  No CPU compute fallback:
// without a CUDA device every entry point fails with PYDSB_ERR_NO_DEVICE.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/pydsb.h"
#include "feas_kernel.cuh"

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

struct Model {
    int device = 0;
    ExecImage img{};
    void* d_code = nullptr;
    void* d_consts = nullptr;
    void* d_tabs = nullptr;
    std::vector<uint32_t> xpt;          // pre-decoded point tape (16 words per instruction)
    unsigned long long* d_counters = nullptr;
    double* d_partial = nullptr;
    size_t partial_cap = 0;
    int num_sms = 0, blocks_per_sm = 0, regs = 0;
    unsigned long long launches = 0;
    // pinned + device staging for the *_host entry points
    double* h_pin = nullptr;
    size_t h_pin_cap = 0;
    double* d_stage = nullptr;
    size_t d_stage_cap = 0;
    cudaStream_t own_stream = nullptr;
};

const Section* find(const std::vector<Section>& secs, const char* tag)
{
    for (const auto& s : secs)
        if (strncmp(s.tag, tag, 4) == 0) return &s;
    return nullptr;
}

typedef void (*KernelFn)(const ExecImage, const LaunchArgs);

KernelFn kernel_for(int ns)
{
    switch (ns) {
    case 1: return feas_kernel<1>;
    case 2: return feas_kernel<2>;
    case 3: return feas_kernel<3>;
    default: return nullptr;
    }
}

int align_up(int v, int a) { return (v + a - 1) / a * a; }

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
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PYDSB_ERR_NO_DEVICE, "no CUDA device visible: pyds_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(PYDSB_ERR_INVALID, "device index out of range");
    const unsigned char* p = static_cast<const unsigned char*>(blob);
    uint32_t magic, ver, nsec;
    memcpy(&magic, p, 4); memcpy(&ver, p + 4, 4); memcpy(&nsec, p + 8, 4);
    if (magic != 0x42534450u || ver != 1u) return fail(PYDSB_ERR_INVALID, "bad blob magic/version");
    std::vector<Section> secs;
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
        secs.push_back(s);
    }
    const char* need[] = {"HEAD", "DBLS", "CNST", "ADOT", "BWTS", "BIGM", "ZFIX", "TPRO", "TELM", "TPNT",
                          "TGEE", "RPAR", "RX0_", "RQ0_", "RZ__", "RF__", "RJ__", "RFQ_", "RG__", "CMAP"};
    for (const char* t : need)
        if (!find(secs, t)) return fail(PYDSB_ERR_INVALID, std::string("blob lacks section ") + t);
    const Section* head = find(secs, "HEAD");
    if (head->count < 18) return fail(PYDSB_ERR_INVALID, "HEAD too short");
    const uint32_t* H = reinterpret_cast<const uint32_t*>(head->data);

    Model* m = new Model();
    m->device = device;
    ExecImage& g = m->img;
    g.ns = H[0]; g.nq = H[1]; g.ng = H[2]; g.d_dim = H[3]; g.p_dim = H[4]; g.nz = H[5]; g.nfe = H[6];
    const int ncp = H[7];
    g.nce = H[8]; g.n_units = H[9]; g.ctrl_slot = H[10]; g.xe_slot = H[11]; g.qe_slot = H[12];
    g.x_slot = H[13]; g.max_it = H[14]; g.j_mask = H[15]; g.fq_state_dep = H[16]; g.predictor = H[17];
    if (ncp != 3) { delete m; return fail(PYDSB_ERR_UNSUPPORTED, "kernels are built for ncp = 3 (Radau)"); }
    if (!kernel_for(g.ns)) {
        delete m;
        return fail(PYDSB_ERR_UNSUPPORTED, "register-resident Newton kernels cover 1..3 dynamic states");
    }
    if (g.ng < 1 || g.nfe < 1 || g.p_dim < 1 || g.d_dim < 0) { delete m; return fail(PYDSB_ERR_INVALID, "bad dimensions"); }
    if (g.nq > 8) { delete m; return fail(PYDSB_ERR_UNSUPPORTED, "at most 8 quadrature states"); }
    {
        // f / J / fq outputs of the point tape -> (offset, stride) tables read by the compiled Newton code
        auto resolve = [&](const char* tag, int n, unsigned* off, unsigned* str, unsigned* mask) -> bool {
            const Section* s = find(secs, tag);
            if ((int)s->count < n) return false;
            const uint16_t* v = reinterpret_cast<const uint16_t*>(s->data);
            *mask = 0;
            for (int i = 0; i < n; ++i) {
                off[i] = 0; str[i] = 0;
                if (v[i] == REF_NONE) continue;
                if (v[i] & REF_CONST) return false;
                *mask |= 1u << i;
                off[i] = (v[i] & REF_MASK) * BLOCK * 8u;
                str[i] = (v[i] & REF_WIDE) ? BLOCK * 8u : 0u;
            }
            return true;
        };
        unsigned fmask = 0, jmask = 0;
        if (!resolve("RF__", g.ns, g.f_off, g.f_str, &fmask) || !resolve("RJ__", g.ns * g.ns, g.j_off, g.j_str, &jmask) ||
            !resolve("RFQ_", g.nq, g.fq_off, g.fq_str, &g.fq_mask)) {
            delete m;
            return fail(PYDSB_ERR_INVALID, "bad f/J/fq reference table");
        }
        g.j_mask = jmask;
        // a structurally zero right-hand side reads the always-zero constant slot: point it at offset of X (harmless)
        // and rely on h*0: keep it simple by requiring f to be non-zero
        if (fmask != (1u << g.ns) - 1u) { delete m; return fail(PYDSB_ERR_UNSUPPORTED, "a dynamic state has an identically zero right-hand side"); }
    }
    {
        const double rtol = reinterpret_cast<const double*>(find(secs, "DBLS")->data)[0];
        g.rtol2 = rtol * rtol;
    }
    const Section* adot = find(secs, "ADOT");
    const Section* bw = find(secs, "BWTS");
    if (adot->count != 16 || bw->count != 3) { delete m; return fail(PYDSB_ERR_INVALID, "bad collocation tables"); }
    memcpy(g.adot, adot->data, 16 * sizeof(double));
    memcpy(g.bw, bw->data, 3 * sizeof(double));
    {
        // extrapolation weights of the previous element's cubic: l_k(1 + tau_j), nodes tau from ADOT's
        // generator (Radau-3: 0, (4 -+ sqrt 6)/10, 1)
        const long double s6 = sqrtl(6.0L);
        const long double tau[4] = {0.0L, (4.0L - s6) / 10.0L, (4.0L + s6) / 10.0L, 1.0L};
        for (int j = 0; j < 3; ++j)
            for (int k = 0; k < 4; ++k) {
                long double v = 1.0L;
                for (int l = 0; l < 4; ++l)
                    if (l != k) v *= (1.0L + tau[j + 1] - tau[l]) / (tau[k] - tau[l]);
                g.ext[4 * j + k] = (double)v;
            }
    }

    const Section *tp = find(secs, "TPRO"), *te = find(secs, "TELM"), *tn = find(secs, "TPNT"), *tg = find(secs, "TGEE");
    g.n_pro = tp->count; g.n_elem = te->count; g.n_pt = tn->count; g.n_g = tg->count;
    std::vector<uint64_t> code;
    for (const Section* s : {tp, te, tg}) {
        const uint64_t* c = reinterpret_cast<const uint64_t*>(s->data);
        code.insert(code.end(), c, c + s->count);
    }
    for (uint64_t w : code)
        if ((w & 0xFF) >= OP_COUNT) { delete m; return fail(PYDSB_ERR_INVALID, "unknown opcode in tape"); }
    // pre-decode the point tape: per operand the byte offsets of the three collocation points
    if (g.n_pt > MAX_PT) { delete m; return fail(PYDSB_ERR_UNSUPPORTED, "point tape longer than the constant-bank window"); }
    {
        const uint64_t* c = reinterpret_cast<const uint64_t*>(tn->data);
        m->xpt.assign((size_t)16 * (g.n_pt > 0 ? g.n_pt : 1), 0u);
        for (int i = 0; i < g.n_pt; ++i) {
            const uint64_t w = c[i];
            const unsigned op = (unsigned)(w & 0xFF);
            const unsigned refs[4] = {(unsigned)(w >> 8) & 0x3FFFu, (unsigned)(w >> 22) & 0x3FFFu,
                                      (unsigned)(w >> 36) & 0x3FFFu, (unsigned)(w >> 50) & 0x3FFFu};
            if (op >= OP_BC3) { delete m; return fail(PYDSB_ERR_INVALID, "opcode not allowed in the point tape"); }
            const int nops = (op == OP_FMA || op == OP_FMS || op == OP_FNMA) ? 3
                             : (op == OP_ADD || op == OP_SUB || op == OP_MUL || op == OP_DIV || op == OP_POW ||
                                op == OP_MIN || op == OP_MAX) ? 2 : 1;
            uint32_t* x = &m->xpt[(size_t)16 * i];
            static const int xmap[OP_COUNT] = {/*MOV*/ XOP_MOV, /*NEG*/ XOP_NEG, /*ADD*/ XOP_ADD, /*SUB*/ XOP_SUB, /*MUL*/ XOP_MUL,
                                               /*DIV*/ XOP_DIV, /*FMA*/ XOP_FMA, /*FMS*/ XOP_FMS, /*FNMA*/ XOP_FNMA, /*SQR*/ XOP_SQR,
                                               /*RCP*/ XOP_RCP, /*EXP*/ XOP_COLD, /*LOG*/ XOP_COLD, /*SQRT*/ XOP_COLD, /*SIN*/ XOP_COLD,
                                               /*COS*/ XOP_COLD, /*TANH*/ XOP_COLD, /*POW*/ XOP_COLD, /*ABS*/ XOP_ABS, /*MIN*/ XOP_MIN,
                                               /*MAX*/ XOP_MAX, /*BC3*/ XOP_NOP};
            x[0] = (uint32_t)xmap[op] | (op << 8);
            const bool cold_unary = (xmap[op] == XOP_COLD && nops == 1);
            for (int k = 0; k <= nops; ++k) {
                const unsigned r = refs[k];
                if (r & REF_CONST) { delete m; return fail(PYDSB_ERR_INVALID, "constant operand in the point tape"); }
                if ((int)(r & REF_MASK) + 2 >= g.n_units + 2 && (r & REF_WIDE)) { delete m; return fail(PYDSB_ERR_INVALID, "slot out of range"); }
                for (int p3 = 0; p3 < 3; ++p3)
                    x[1 + 3 * k + p3] = ((r & REF_MASK) + ((r & REF_WIDE) ? p3 : 0)) * BLOCK * 8u;
            }
            if (cold_unary)
                for (int p3 = 0; p3 < 3; ++p3) x[7 + p3] = x[4 + p3];
            if (!(refs[0] & REF_WIDE)) { delete m; return fail(PYDSB_ERR_INVALID, "point tape writes a narrow slot"); }
        }
    }
    const Section* cn = find(secs, "CNST");
    g.n_const = cn->count;

    std::vector<uint16_t> tabs;
    auto push = [&](const char* tag, size_t n) -> bool {
        const Section* s = find(secs, tag);
        if (s->count < n) return false;
        const uint16_t* v = reinterpret_cast<const uint16_t*>(s->data);
        tabs.insert(tabs.end(), v, v + n);
        return true;
    };
    bool ok = push("RPAR", (size_t)g.d_dim + g.p_dim) && push("RX0_", g.ns) && push("RQ0_", g.nq) && push("RZ__", g.nz) &&
              push("ZFIX", g.nz) && push("RG__", g.ng) && push("CMAP", (size_t)g.nfe * g.nce);
    if (!ok) { delete m; return fail(PYDSB_ERR_INVALID, "reference table shorter than its dimension"); }
    g.n_tabs = (int)tabs.size();

    // shared-memory carve-up
    int o = 128;                                            // mbarriers + reduction scratch
    g.off_code = o; o += (int)code.size() * 8;
    g.off_const = o; o += (g.n_const + 1) * 8;
    g.off_tabs = o; o = align_up(o + g.n_tabs * 2, 128);
    g.off_land = o; o = align_up(o + BLOCK * g.p_dim * 8, 128);
    g.off_slots = o; o += g.n_units * BLOCK * 8;
    g.smem_bytes = o;

    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_code, code.size() * 8 + 8);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_code, code.data(), code.size() * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_consts, (size_t)g.n_const * 8 + 8);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_consts, cn->data, (size_t)g.n_const * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_tabs, tabs.size() * 2 + 8);
    if (e == cudaSuccess) e = cudaMemcpy(m->d_tabs, tabs.data(), tabs.size() * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&m->d_counters, PYDSB_CNT_COUNT * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(m->d_counters, 0, PYDSB_CNT_COUNT * sizeof(unsigned long long));
    cudaDeviceProp prop{};
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess && (size_t)g.smem_bytes > prop.sharedMemPerBlockOptin) {
        pydsb_model_destroy(m);
        return fail(PYDSB_ERR_UNSUPPORTED, "model needs " + std::to_string(g.smem_bytes) +
                                               " B of shared memory per CTA, device offers " +
                                               std::to_string(prop.sharedMemPerBlockOptin));
    }
    KernelFn fn = kernel_for(g.ns);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, g.smem_bytes);
    int bps = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, fn, BLOCK, g.smem_bytes);
    cudaFuncAttributes fa{};
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, fn);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        std::string msg = std::string("model upload: ") + cudaGetErrorString(e);
        pydsb_model_destroy(m);
        return fail(PYDSB_ERR_CUDA, msg);
    }
    g.code = static_cast<const uint64_t*>(m->d_code);
    g.consts = static_cast<const double*>(m->d_consts);
    g.tabs = static_cast<const uint16_t*>(m->d_tabs);
    m->num_sms = prop.multiProcessorCount;
    m->blocks_per_sm = bps > 0 ? bps : 1;
    m->regs = fa.numRegs;
    *handle = m;
    return PYDSB_OK;
}

int pydsb_model_destroy(void* handle)
{
    Model* m = static_cast<Model*>(handle);
    if (!m) return PYDSB_OK;
    cudaSetDevice(m->device);
    if (m->d_code) cudaFree(m->d_code);
    if (m->d_consts) cudaFree(m->d_consts);
    if (m->d_tabs) cudaFree(m->d_tabs);
    if (m->d_counters) cudaFree(m->d_counters);
    if (m->d_partial) cudaFree(m->d_partial);
    if (m->d_stage) cudaFree(m->d_stage);
    if (m->h_pin) cudaFreeHost(m->h_pin);
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    delete m;
    return PYDSB_OK;
}

int pydsb_model_info(void* handle, long long* out, int n)
{
    Model* m = static_cast<Model*>(handle);
    if (!m || !out) return fail(PYDSB_ERR_INVALID, "null handle/out");
    long long v[PYDSB_INFO_COUNT] = {m->img.ns, m->img.nq, m->img.ng, m->img.d_dim, m->img.p_dim, m->img.nz,
                                     m->img.nfe, 3, m->img.n_units, m->img.smem_bytes, m->blocks_per_sm,
                                     m->num_sms, m->regs};
    for (int i = 0; i < n && i < PYDSB_INFO_COUNT; ++i) out[i] = v[i];
    return PYDSB_OK;
}

int pydsb_eval(void* handle, const double* d, long long n_d, const double* theta, const double* w, long long n_t,
               const double* z, int z_mode, double* g_out, double* ind_out, double* P_out, void* stream)
{
    Model* m = static_cast<Model*>(handle);
    if (!m) return fail(PYDSB_ERR_INVALID, "null handle");
    if (n_d < 0 || n_t < 0) return fail(PYDSB_ERR_INVALID, "negative size");
    if (n_d == 0) return PYDSB_OK;
    if (!theta || (!d && m->img.d_dim > 0)) return fail(PYDSB_ERR_INVALID, "null d/theta");
    if (z_mode < 0 || z_mode > 3) return fail(PYDSB_ERR_INVALID, "bad z_mode");
    if (z_mode != 0 && !z && m->img.nz > 0) return fail(PYDSB_ERR_INVALID, "z_mode needs a z array");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CUDA_TRY(cudaSetDevice(m->device));
    if (n_t == 0) {                     // empty theta set: P = 0 for every design point
        if (P_out) CUDA_TRY(cudaMemsetAsync(P_out, 0, (size_t)n_d * sizeof(double), st));
        return PYDSB_OK;
    }
    const long long n_tiles_ll = (n_t + BLOCK - 1) / BLOCK;
    if (n_tiles_ll > 0x7fffffffLL) return fail(PYDSB_ERR_INVALID, "n_t too large");
    const int n_tiles = (int)n_tiles_ll;
    const long long total = n_d * n_tiles_ll;
    int rc = ensure_partial(m, (size_t)total);
    if (rc) return rc;

    LaunchArgs la{};
    la.d = d; la.theta = theta; la.w = w; la.z = z; la.z_mode = z_mode;
    la.n_d = n_d; la.n_t = n_t; la.n_tiles = n_tiles;
    la.theta_tma_ok = (reinterpret_cast<uintptr_t>(theta) % 16 == 0) ? 1 : 0;
    la.partial = m->d_partial; la.g_out = g_out; la.ind_out = ind_out; la.counters = m->d_counters;

    const long long max_ctas = (long long)m->num_sms * m->blocks_per_sm;
    const int grid = (int)(total < max_ctas ? total : max_ctas);
    KernelFn fn = kernel_for(m->img.ns);
    // several models may share one template instantiation: (re)assert this model's smem size and
    // (re)load its point tape into the constant bank (stream ordered)
    CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, m->img.smem_bytes));
    CUDA_TRY(cudaMemcpyToSymbolAsync(c_pt, m->xpt.data(), m->xpt.size() * sizeof(uint32_t), 0, cudaMemcpyHostToDevice, st));
    fn<<<grid, BLOCK, m->img.smem_bytes, st>>>(m->img, la);
    CUDA_TRY(cudaGetLastError());
    m->launches += 1;
    if (P_out) {
        const int tb = 128;
        reduce_partials_kernel<<<(unsigned)((n_d + tb - 1) / tb), tb, 0, st>>>(m->d_partial, n_tiles, n_d, P_out);
        CUDA_TRY(cudaGetLastError());
        m->launches += 1;
    }
    return PYDSB_OK;
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
    const ExecImage& g = m->img;
    const size_t nd = (size_t)n_d * g.d_dim, nth = (size_t)n_t * g.p_dim, nw = w ? (size_t)n_t : 0;
    size_t nz = 0;
    if (z && z_mode == 1) nz = g.nz;
    if (z && z_mode == 2) nz = (size_t)n_d * g.nz;
    if (z && z_mode == 3) nz = (size_t)n_d * n_t * g.nz;
    const size_t n_in = nd + nth + nw + nz;
    const size_t ng = g_out ? (size_t)n_d * n_t * g.ng : 0, ni = ind_out ? (size_t)n_d * n_t : 0, np = P_out ? (size_t)n_d : 0;
    const size_t n_out = ng + ni + np;
    // layout (doubles): inputs padded so that theta starts 16-byte aligned
    const size_t o_d = 0, o_th = (nd + 1) / 2 * 2, o_w = o_th + nth, o_z = o_w + nw, o_out = (o_z + nz + 1) / 2 * 2;
    const size_t need = o_out + n_out;
    (void)n_in;
    if (need > m->h_pin_cap) {
        if (m->h_pin) cudaFreeHost(m->h_pin);
        if (m->d_stage) cudaFree(m->d_stage);
        m->h_pin = nullptr; m->d_stage = nullptr; m->h_pin_cap = m->d_stage_cap = 0;
        CUDA_TRY(cudaMallocHost(&m->h_pin, need * sizeof(double)));
        CUDA_TRY(cudaMalloc(&m->d_stage, need * sizeof(double)));
        m->h_pin_cap = m->d_stage_cap = need;
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
