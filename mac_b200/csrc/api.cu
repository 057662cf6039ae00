// libmacb200.so -- C-ABI (include/macb200.h) and host-side drivers over the kernels in kernels.cuh.
//
// One handle = one graph (fixed + candidate edges) resident on one B200, one stream, one set of
// captured CUDA graphs.  The Frank-Wolfe loop (frankwolfe.py:10-79 as called from mac.py:196) runs
// entirely on the device; the host only (a) solves the tiny tridiagonal Rayleigh-Ritz problem of the
// Lanczos iteration and (b) evaluates the two scalar stopping tests per FW iteration.
#include "../../include/macb200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include <chrono>
#include <thread>
#include <dlfcn.h>

#include "kernels.cuh"
#include "tridiag.h"

using namespace macb;

namespace {

thread_local std::string g_create_error;

// The persisting-L2 carve-out is a property of the device, not of a handle: keep the largest carve any live handle asked
// for and hand it back only when the last pinning handle on that device goes away.
std::mutex g_l2_mutex;
struct L2Pin { int refs = 0; size_t carve = 0; };
std::map<int, L2Pin> g_l2_pins;

constexpr int kGraphSteps = 32;            // granularity of the Lanczos basis capacity
constexpr size_t kFlushBytes = 512u << 20; // > 126 MB L2

struct CudaFail {
    cudaError_t e;
    const char* what;
    int line;
};

#define CK(call)                                              \
    do {                                                      \
        cudaError_t _e = (call);                              \
        if (_e != cudaSuccess) throw CudaFail{_e, #call, __LINE__}; \
    } while (0)

struct ArgFail {
    std::string msg;
    int code;
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the FUNCTION, shared by every handle in the process: only ever
// raise it (a handle with a small graph must not lower it under a handle with a large one -- the cooperative launch of the
// large one then fails with "too many blocks").
void raise_dyn_smem(const void* fn, size_t bytes) {
    static std::mutex mu;
    static std::map<const void*, size_t> cur;
    std::lock_guard<std::mutex> lk(mu);
    size_t& c = cur[fn];
    if (bytes <= c) return;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    c = bytes;
}

template <typename T>
T* dalloc(size_t count) {
    T* p = nullptr;
    if (count == 0) count = 1;
    CK(cudaMalloc(&p, count * sizeof(T)));
    return p;
}

}  // namespace

struct macb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    int grid_max = 148 * 8;
    int W = 8;  // lanes per row in the row-parallel kernels

    int32_t n = 0;
    int ld = 0;
    int64_t nf = 0, m = 0, nnz = 0;

    int *d_rp = nullptr, *d_col = nullptr, *d_eid = nullptr;
    double *d_val = nullptr, *d_diag = nullptr, *d_ew = nullptr;
    int *d_ci = nullptr, *d_cj = nullptr;
    double *d_kappa = nullptr, *d_x = nullptr, *d_g = nullptr, *d_tmp_m = nullptr;
    uint8_t* d_sel = nullptr;
    double *d_v = nullptr, *d_y = nullptr, *d_x0 = nullptr, *d_tmp_n = nullptr;
    double x0_norm = 0.0;

    // Lanczos
    double* d_basis = nullptr;
    int64_t basis_cap = 0;  // number of Lanczos steps the basis can hold (cap + 1 vectors)
    double *d_alpha = nullptr, *d_beta = nullptr, *d_coef = nullptr;
    LzScalars* d_sc = nullptr;
    double *h_alpha = nullptr, *h_beta = nullptr;  // pinned
    LzScalars* h_sc = nullptr;                     // pinned
    // persistent engine
    int persist_v = 3;             // 3: slot-parallel kernel (k_lanczos_slots); 1: row-parallel (k_lanczos_persist)
    int *d_chunk_ptr = nullptr, *d_chunk_row = nullptr;
    size_t slots_smem = 0;
    int slots_cache_cols = 0, slots_prod_cap = 0;
    // chunked jagged-diagonal SpMV (k_spmv_jds): built on demand by macb_spmv_engine(h, 1)
    int spmv_engine = 0;           // 0: k_spmv (CSR, W lanes per row), 1: k_spmv_jds
    int sj_nchunks = 0;
    int *d_sj_chunk_row = nullptr, *d_sj_chunk_jd = nullptr, *d_sj_jd = nullptr, *d_sj_perm = nullptr, *d_sj_len = nullptr,
        *d_sj_eid = nullptr, *d_sj_col0 = nullptr, *d_sj_col = nullptr;
    unsigned int* d_sj_word = nullptr;
    bool sj_col16 = false;
    int64_t* d_sj_chunk_slot = nullptr;
    double* d_sj_val = nullptr;
    int vec_batch = 5;             // gathers in flight per thread in k_lanczos_pipe (3..8, chosen from the slots per thread)
    bool pipe = true;              // k_lanczos_pipe is built (pipelined recurrence, reduction off the critical path)
    size_t pipe_smem = 0;
    // Two sets of read-back buffers / events, so that iteration i+1 of the Frank-Wolfe loop can be enqueued before the host has
    // looked at the scalars of iteration i (set 0 = the members above; use_slot() points the members at a set)
    struct IterSlot {
        LzScalars* h_sc = nullptr;
        RrOut* h_rr = nullptr;
        SelState* h_sel = nullptr;
        cudaEvent_t done = nullptr, it0 = nullptr, it1 = nullptr, lz0 = nullptr, lz1 = nullptr;
    } slot[2];
    bool slots_ready = false;
    double* d_x_alt = nullptr;     // the other iterate buffer of the pipelined Frank-Wolfe loop
    bool force_host_rr = false;    // this solve only: host-driven path (fallback after a failed device-side decision)
    bool dev_rr = false;           // Rayleigh-Ritz / stop decision on the device (extra CTA of the k_lanczos_pipe launch)
    double *d_rr_a = nullptr, *d_rr_b = nullptr, *d_rr_b2 = nullptr, *d_rr_binv = nullptr, *d_rr_s = nullptr, *d_rr_w = nullptr;
    RrOut* d_rr_out = nullptr;
    RrOut* h_rr = nullptr;         // pinned
    RrArgs rr_launch = {};         // filled by enqueue_fiedler_device for the next k_lanczos_pipe launch (enabled = 0 otherwise)
    int* d_dev_stop = nullptr;
    double lz_algo_bytes = 0.0;    // bench mode: sum over timed launches of phases x algorithmic bytes at that launch's support
    int64_t c_dev_fallbacks = 0;   // eigen-solves that fell back from the device-side decision to the host-driven path
    bool sect_joint = false;       // d_sect[1] points into d_sect[0]'s allocation
    double* d_zprev = nullptr;     // z_{j-1} across launches of k_lanczos_pipe
    bool l2_pinned = false;        // this handle holds a reference on the device's persisting-L2 carve-out
    double* h_ab = nullptr;        // host-mapped [2 * (cap + 2)]
    int* h_stop = nullptr;         // host-mapped stop flag (the Lanczos kernels sample it once per phase)
    int check_div = 8;             // Rayleigh-Ritz check interval = k / check_div (24 when the graph fills the GPU)
    int ab_dirty = 0;              // h_ab entries [0, ab_dirty) may hold values of an earlier launch
    int p_ncta = 1;
    int* d_row_start = nullptr;
    double* d_sect[2] = {nullptr, nullptr};
    LzPartRec* d_precs = nullptr;
    LzPersistState* d_pst = nullptr;
    long long* d_ptiming = nullptr;
    std::vector<int32_t> h_rp;  // host copy of row_ptr (row partition)
    std::vector<int32_t> h_col, h_eid;  // host copies of the pattern until the persistent engine has been set up
    // sliced layout of the pipelined kernel (k_lanczos_pipe, persist_v == 5)
    int *d_jrow = nullptr, *d_jlen = nullptr, *d_jcol = nullptr, *d_jeid = nullptr, *d_jd = nullptr;
    double* d_jval = nullptr;
    double* d_xrec = nullptr;      // inboxes of the all-to-all record exchange of k_lanczos_pipe
    int pipe_pos_cap = 0, pipe_slot_cap = 0;   // k_lanczos_pipe: product positions / slots per CTA (shared-memory sizes)

    // reductions / selection
    double* d_partials = nullptr;
    unsigned int* d_counter = nullptr;
    SelState* d_sel_state = nullptr;
    unsigned int* d_blockcnt = nullptr;
    SelState* h_sel_state = nullptr;  // pinned
    SelState* d_sel_state2 = nullptr;
    unsigned int* d_sel2_hist = nullptr;       // two-pass top-k: global 15-bit histogram, state, candidate keys of the chosen bin
    Sel2State* d_sel2 = nullptr;
    unsigned long long* d_sel_cand = nullptr;
    double *d_tmp_m2 = nullptr, *d_tmp_m3 = nullptr;
    cudaGraphExec_t sel_graph = nullptr;

    void* d_flush = nullptr;

    bool have_x = false, have_v = false, have_g = false, have_sel = false;
    bool have_prev_v = false;  // d_v holds a (possibly stale) Fiedler vector usable as a warm start
    double lnorm = 0.0;
    int64_t nnz_active = 0;
    double min_sel_tol = 1e-10;

    int64_t c_launches = 0, c_spmv = 0, c_steps = 0, c_solves = 0;
    double lz_kernel_ms = 0.0;     // CUDA-event time of the Lanczos kernels (bench mode only)
    int64_t lz_kernel_phases = 0;
    cudaEvent_t lz0 = nullptr, lz1 = nullptr;
    double phase_ms[MACB_T_COUNT] = {0, 0, 0, 0, 0, 0};
    bool profile = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool bench_time_iters = false, bench_flush = false;
    cudaEvent_t it0 = nullptr, it1 = nullptr;
    std::vector<double> iter_ms;

    std::string err;

    ReduceWS ws() const { return ReduceWS{d_partials, d_counter}; }
    int grid_for(int64_t work_items) const {
        int64_t b = (work_items + kBlock - 1) / kBlock;
        if (b < 1) b = 1;
        return (int)std::min<int64_t>(b, grid_max);
    }
    int grid_rows() const { return grid_for((int64_t)n * W); }
};

namespace {

struct PhaseTimer {
    macb_ctx* c;
    int phase;
    PhaseTimer(macb_ctx* c_, int p) : c(c_), phase(p) {
        if (c->profile) cudaEventRecord(c->ev0, c->stream);
    }
    ~PhaseTimer() {
        if (c->profile) {
            cudaEventRecord(c->ev1, c->stream);
            cudaEventSynchronize(c->ev1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, c->ev0, c->ev1);
            c->phase_ms[phase] += ms;
        }
    }
};

int pick_width(double avg_row) {
    const char* env = getenv("MACB_SPMV_W");
    if (env) {
        int w = atoi(env);
        if (w == 2 || w == 4 || w == 8 || w == 16 || w == 32) return w;
    }
    if (avg_row <= 3.0) return 2;
    if (avg_row <= 8.0) return 4;
    if (avg_row <= 28.0) return 8;
    if (avg_row <= 64.0) return 16;
    return 32;
}

#define DISPATCH_W(W_, ...)                   \
    switch (W_) {                             \
        case 2: { constexpr int WW = 2; __VA_ARGS__; } break;   \
        case 4: { constexpr int WW = 4; __VA_ARGS__; } break;   \
        case 8: { constexpr int WW = 8; __VA_ARGS__; } break;   \
        case 16: { constexpr int WW = 16; __VA_ARGS__; } break; \
        default: { constexpr int WW = 32; __VA_ARGS__; } break; \
    }

// ------------------------------------------------------------------------------------------------
// Host pattern builder: union pattern, off-diagonals only, rows sorted by (column, edge id).
void build_pattern(int32_t n, int64_t nf, const int32_t* fi, const int32_t* fj, int64_t m, const int32_t* ci,
                   const int32_t* cj, std::vector<int32_t>& rp, std::vector<int32_t>& col, std::vector<int32_t>& eid) {
    const int64_t ne = nf + m;
    auto EI = [&](int64_t e) { return e < nf ? fi[e] : ci[e - nf]; };
    auto EJ = [&](int64_t e) { return e < nf ? fj[e] : cj[e - nf]; };
    std::vector<int64_t> cnt((size_t)n + 1, 0);
    int64_t nnz = 0;
    for (int64_t e = 0; e < ne; ++e) {
        int32_t a = EI(e), b = EJ(e);
        if (a < 0 || a >= n || b < 0 || b >= n) throw ArgFail{"edge endpoint out of range [0, num_nodes)", MACB_ERR_ARG};
        if (a == b) continue;
        cnt[a + 1]++;
        cnt[b + 1]++;
        nnz += 2;
    }
    if (nnz >= (int64_t)std::numeric_limits<int32_t>::max()) throw ArgFail{"pattern exceeds int32 slots", MACB_ERR_ARG};
    // pass 1: bucket by column (so that the stable pass 2 leaves rows sorted by column, then edge id)
    std::vector<int64_t> cstart((size_t)n + 1, 0);
    for (int32_t i = 0; i < n; ++i) cstart[i + 1] = cstart[i] + cnt[i + 1];  // symmetric: column counts == row counts
    std::vector<int32_t> t_row((size_t)nnz), t_eid((size_t)nnz);
    {
        std::vector<int64_t> pos(cstart.begin(), cstart.end() - 1);
        for (int64_t e = 0; e < ne; ++e) {
            int32_t a = EI(e), b = EJ(e);
            if (a == b) continue;
            // entry (row a, col b) goes to column bucket b; entry (row b, col a) to bucket a
            int64_t p = pos[b]++;
            t_row[p] = a;
            t_eid[p] = (int32_t)e;
            p = pos[a]++;
            t_row[p] = b;
            t_eid[p] = (int32_t)e;
        }
    }
    rp.assign((size_t)n + 1, 0);
    for (int32_t i = 0; i < n; ++i) rp[i + 1] = (int32_t)(rp[i] + cnt[i + 1]);
    col.resize((size_t)nnz);
    eid.resize((size_t)nnz);
    {
        std::vector<int64_t> pos((size_t)n);
        for (int32_t i = 0; i < n; ++i) pos[i] = rp[i];
        for (int32_t c = 0; c < n; ++c)
            for (int64_t p = cstart[c]; p < cstart[c + 1]; ++p) {
                int64_t q = pos[t_row[p]]++;
                col[q] = c;
                eid[q] = t_eid[p];
            }
    }
}

void free_all(macb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->l2_pinned) {   // the weights were pinned in the persisting part of L2 (setup_persist)
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.num_bytes = 0;   // drop this stream's window; other handles' lines stay where they are
        if (c->stream) cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        std::lock_guard<std::mutex> lk(g_l2_mutex);
        L2Pin& pin = g_l2_pins[c->device];
        if (--pin.refs <= 0) {
            cudaCtxResetPersistingL2Cache();
            pin = L2Pin{};
        }
        cudaGetLastError();
    }
    if (c->sel_graph) cudaGraphExecDestroy(c->sel_graph);
    void* dptrs[] = {c->d_rp, c->d_col, c->d_eid, c->d_val, c->d_diag, c->d_ew, c->d_ci, c->d_cj, c->d_kappa,
                     c->d_x, c->d_g, c->d_tmp_m, c->d_sel, c->d_v, c->d_y, c->d_x0, c->d_tmp_n, c->d_basis,
                     c->d_alpha, c->d_beta, c->d_coef, c->d_sc, c->d_partials, c->d_counter,
                     c->d_sel_state, c->d_sel_state2, c->d_tmp_m2, c->d_tmp_m3, c->d_blockcnt, c->d_flush, c->d_row_start, c->d_sect[0], c->sect_joint ? nullptr : c->d_sect[1], c->d_precs, c->d_chunk_ptr, c->d_chunk_row,
                     c->d_pst, c->d_ptiming, c->d_jrow, c->d_jlen, c->d_jcol, c->d_jeid, c->d_jd, c->d_jval, c->d_xrec, c->d_sj_chunk_row, c->d_sj_chunk_jd, c->d_sj_jd,
                     c->d_sj_perm, c->d_sj_len, c->d_sj_col, c->d_sj_eid, c->d_sj_chunk_slot, c->d_sj_word, c->d_sj_val, c->d_sj_col0, c->d_zprev, c->d_rr_a, c->d_rr_b, c->d_rr_b2, c->d_rr_binv, c->d_rr_s, c->d_rr_out, c->d_dev_stop, c->d_rr_w, c->d_sel2_hist, c->d_sel2, c->d_sel_cand};
    for (void* p : dptrs)
        if (p) cudaFree(p);
    if (c->h_alpha) cudaFreeHost(c->h_alpha);
    if (c->h_beta) cudaFreeHost(c->h_beta);
    if (c->h_sc) cudaFreeHost(c->h_sc);
    if (c->h_sel_state) cudaFreeHost(c->h_sel_state);
    if (c->h_ab) cudaFreeHost(c->h_ab);
    if (c->h_stop) cudaFreeHost(c->h_stop);
    if (c->slots_ready) {
        c->h_sc = c->slot[0].h_sc; c->h_rr = c->slot[0].h_rr; c->h_sel_state = c->slot[0].h_sel;
        c->it0 = c->slot[0].it0; c->it1 = c->slot[0].it1; c->lz0 = c->slot[0].lz0; c->lz1 = c->slot[0].lz1;
        if (c->slot[1].h_sc) cudaFreeHost(c->slot[1].h_sc);
        if (c->slot[1].h_rr) cudaFreeHost(c->slot[1].h_rr);
        if (c->slot[1].h_sel) cudaFreeHost(c->slot[1].h_sel);
        for (cudaEvent_t e : {c->slot[1].it0, c->slot[1].it1, c->slot[1].lz0, c->slot[1].lz1, c->slot[0].done, c->slot[1].done})
            if (e) cudaEventDestroy(e);
    }
    if (c->d_x_alt) cudaFree(c->d_x_alt);
    if (c->h_rr) cudaFreeHost(c->h_rr);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->it0) cudaEventDestroy(c->it0);
    if (c->it1) cudaEventDestroy(c->it1);
    if (c->lz0) cudaEventDestroy(c->lz0);
    if (c->lz1) cudaEventDestroy(c->lz1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

// Deterministic built-in start vector (used when the caller supplies none): splitmix64 + Box-Muller.
void default_start(int32_t n, std::vector<double>& x0) {
    x0.resize(n);
    uint64_t s = 0x9E3779B97F4A7C15ull * 7ull;
    auto next = [&]() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    for (int32_t i = 0; i < n; ++i) {
        double u1 = ((next() >> 11) + 1.0) * (1.0 / 9007199254740993.0);
        double u2 = (next() >> 11) * (1.0 / 9007199254740992.0);
        x0[i] = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
}

void upload_start(macb_ctx* c, const double* x0_in) {
    std::vector<double> x0(c->n);
    if (x0_in)
        std::copy(x0_in, x0_in + c->n, x0.begin());
    else
        default_start(c->n, x0);
    // project out the all-ones vector (nx:206-210,230) once on the host
    long double s = 0.0L;
    for (double v : x0) s += v;
    double mean = (double)(s / c->n);
    long double q = 0.0L;
    for (double& v : x0) {
        v -= mean;
        q += (long double)v * v;
    }
    c->x0_norm = std::sqrt((double)q);
    if (c->n >= 2 && !(c->x0_norm > 0.0))
        throw ArgFail{"start vector is constant (no component orthogonal to 1)", MACB_ERR_ARG};
    CK(cudaMemcpyAsync(c->d_x0, x0.data(), sizeof(double) * c->n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
}

// ------------------------------------------------------------------------------------------------ launches
void launch_assemble(macb_ctx* c, bool sync = true) {
    PhaseTimer pt(c, MACB_T_ASSEMBLE);
    const int grid = c->grid_rows();
    DISPATCH_W(c->W, k_assemble<WW><<<grid, kBlock, 0, c->stream>>>(c->n, c->d_rp, c->d_eid, c->d_ew, c->d_val,
                                                                       c->d_diag, c->d_sc, c->ws()));
    if (c->persist_v == 5) {
        k_assemble_jds<<<c->grid_for(c->nnz), kBlock, 0, c->stream>>>(c->nnz, c->d_jeid, c->d_ew, c->d_jval);
        c->c_launches++;
    }
    if (c->d_sj_val) {
        k_assemble_jds<<<c->grid_for(c->nnz), kBlock, 0, c->stream>>>(c->nnz, c->d_sj_eid, c->d_ew, c->d_sj_val);
        c->c_launches++;
    }
    CK(cudaGetLastError());
    c->c_launches++;
    c->have_v = false;
    c->have_g = false;
    c->have_sel = false;
    if (!sync) return;   // the caller reads ||L||_inf and the support size with the rest of the iteration's scalars
    CK(cudaMemcpyAsync(c->h_sc, c->d_sc, sizeof(LzScalars), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->lnorm = c->h_sc->lnorm;
    c->nnz_active = c->h_sc->nnz_active;
    c->have_v = false;
    c->have_g = false;
    c->have_sel = false;
}

template <int MODE>
void launch_spmv(macb_ctx* c, const double* x, double* y) {
    SpmvArgs a;
    a.n = c->n;
    a.ld = c->ld;
    a.rp = c->d_rp;
    a.col = c->d_col;
    a.val = c->d_val;
    a.diag = c->d_diag;
    a.x = x;
    a.y = y;
    a.sc = c->d_sc;
    a.ws = c->ws();
    if (MODE == 0 && c->spmv_engine == 1 && c->d_sj_val) {
        SpmvJdsArgs j{c->sj_nchunks, c->d_sj_chunk_row, c->d_sj_chunk_slot, c->d_sj_chunk_jd, c->d_sj_col0, c->d_sj_jd, c->d_sj_perm,
                      c->d_sj_len, c->d_sj_word, c->d_sj_col, c->d_sj_val, c->d_diag, x, y};
        const int grid = std::min(c->sj_nchunks, kSjMinBlocks * c->sm_count);
        if (c->sj_col16) k_spmv_jds<true><<<grid, kSjBlock, 0, c->stream>>>(j);
        else k_spmv_jds<false><<<grid, kSjBlock, 0, c->stream>>>(j);
        CK(cudaGetLastError());
        return;
    }
    const int grid = c->grid_rows();
    DISPATCH_W(c->W, k_spmv<WW, MODE><<<grid, kBlock, 0, c->stream>>>(a));
    CK(cudaGetLastError());
}

template <int W>
void launch_persist_w(macb_ctx* c, LzPersistArgs& a) {
    void* params[] = {&a};
    CK(cudaLaunchCooperativeKernel((void*)k_lanczos_persist<W>, dim3(a.ncta), dim3(kPBlock), params, 0, c->stream));
}

// One cooperative launch = `nphases` fused Lanczos phases (kernels.cuh, k_lanczos_persist).
void launch_persist(macb_ctx* c, int nphases, bool async = false) {
    LzPersistArgs a;
    a.n = c->n;
    a.ld = c->ld;
    a.nphases = nphases;
    a.ncta = c->p_ncta;
    a.rp = c->d_rp;
    a.col = c->d_col;
    a.val = c->d_val;
    a.row_start = c->d_row_start;
    a.sect[0] = c->d_sect[0];
    a.sect[1] = c->d_sect[1];
    a.basis = c->d_basis;
    a.alpha = c->d_alpha;
    a.beta = c->d_beta;
    a.recs = c->d_precs;
    a.st = c->d_pst;
    a.timing = c->d_ptiming;
    a.ab_host = async ? c->h_ab : nullptr;
    a.stop = async ? c->h_stop : nullptr;
    if (c->bench_time_iters) CK(cudaEventRecord(c->lz0, c->stream));
    if (c->persist_v == 4) {
        RrArgs R = c->rr_launch;
        R.smem_doubles = (int)(c->slots_smem / 8);
        if (R.enabled) {
            a.ab_host = nullptr;
            a.stop = nullptr;
        }
        const double* dg = c->d_diag;
        void* sparams[] = {&a, &dg, &R};
        // cooperative: the solver CTA and the Rayleigh-Ritz CTA must be co-resident (the second stops the first)
        CK(cudaLaunchCooperativeKernel((void*)k_lanczos_small2, dim3(R.enabled ? 2 : 1), dim3(kPBlock), sparams, c->slots_smem, c->stream));
    } else if (c->persist_v == 5) {
        LzJdsArgs J{c->d_row_start, c->d_jcol, c->d_jval, c->d_jd, c->pipe_pos_cap, c->pipe_slot_cap, c->d_xrec, c->d_diag, c->d_jrow};
        LzPipeArgs P{c->d_sc, c->d_zprev, c->rr_launch.enabled ? c->d_dev_stop : nullptr};
        RrArgs R = c->rr_launch;
        R.smem_doubles = (int)(c->pipe_smem / 8);
        if (R.enabled) {   // device-side decision: nothing is streamed to, or polled from, the host
            a.ab_host = nullptr;
            a.stop = nullptr;
        }
        void* pparams[] = {&a, &J, &P, &R};
        void* pf;
        switch (c->vec_batch) {
            case 3: pf = (void*)k_lanczos_pipe<3>; break;
            case 4: pf = (void*)k_lanczos_pipe<4>; break;
            case 6: pf = (void*)k_lanczos_pipe<6>; break;
            case 7: pf = (void*)k_lanczos_pipe<7>; break;
            case 8: pf = (void*)k_lanczos_pipe<8>; break;
            default: pf = (void*)k_lanczos_pipe<5>; break;
        }
        CK(cudaLaunchCooperativeKernel(pf, dim3(a.ncta + (R.enabled ? 1 : 0)), dim3(kPBlock), pparams, c->pipe_smem, c->stream));
    } else if (c->persist_v == 3) {
        LzChunkArgs ch{c->d_chunk_ptr, c->d_chunk_row, c->slots_cache_cols, c->slots_prod_cap};
        void* params[] = {&a, &ch};
        CK(cudaLaunchCooperativeKernel((void*)k_lanczos_slots, dim3(a.ncta), dim3(kPBlock), params, c->slots_smem, c->stream));
    } else {
        DISPATCH_W(c->W, launch_persist_w<WW>(c, a));
    }
    if (c->bench_time_iters) CK(cudaEventRecord(c->lz1, c->stream));
    c->c_launches += 1;
    if (!async) {
        c->c_spmv += nphases;
        c->c_steps += nphases;
    }
}

// Sliced layout of the Lanczos kernel k_lanczos_pipe (pure host code, also exported as macb_host_build_slices for the CPU
// test-suite).  CTA b owns the rows row_start[b] .. row_start[b+1] and their slots rp[row_start[b]] .. rp[row_start[b+1]].
// Inside a CTA the rows are renumbered by decreasing length ("engine numbering", jrow[engine row] = caller row, jlen = its
// length); warp w of the CTA owns the engine rows 32 w .. 32 w + 31 -- one SLICE.  A slice is stored like a small ELLPACK
// matrix padded to its longest row (its first, L_w entries) with a lane stride of kLzSlice = 33 doubles:
//     position of entry d of engine row t = base_w + d * kLzSlice + (t & 31),      base_w = sum_{w' < w} L_w' * kLzSlice
// so that the row sums of the kernel need no table of diagonal starts (the jagged-diagonal predecessor paid a dependent
// shared-memory round trip per eight products for it, measured: 2 400 of a step's 12 500 cycles) and run unpredicated: the
// padding positions are never written and stay zero.  The rows being sorted, the padding is ~4 %, the odd stride 3 %.
// jw[b * kLzSliceTab + 2 w] = base_w, [2 w + 1] = L_w;  positions[b] = the CTA's position count.
// The CTA's SLOTS are stored in column order and carry their position: jcol = column | position << 17 (n < 2^17 - 1, which
// leaves the all-ones column for "inactive", and < 2^15 positions per CTA).  Which entry d of its row a slot uses is free;
// with bankfit it is chosen so that the 16 positions a half-warp scatters to fall into different 8-byte banks -- the odd
// stride is what lets d move the bank.
void build_slice_layout(int n, const int32_t* rp, const int32_t* col, const int32_t* eid, int ncta, const int* row_start, bool bankfit,
                        int* jrow, int* jlen, int* jcol, int* jeid, int* jw, int* positions) {
    std::vector<int> order, inv((size_t)n);   // inv[caller id] = engine id
    for (int b = 0; b < ncta; ++b) {
        const int ra = row_start[b], R = row_start[b + 1] - ra;
        order.resize(R);
        for (int t = 0; t < R; ++t) order[t] = ra + t;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return rp[x + 1] - rp[x] > rp[y + 1] - rp[y]; });
        for (int t = 0; t < R; ++t) {
            jrow[ra + t] = order[t];
            jlen[ra + t] = rp[order[t] + 1] - rp[order[t]];
            inv[order[t]] = ra + t;
        }
    }
    struct Slot { int col, t, eid; };
    std::vector<Slot> slots;
    std::vector<std::vector<unsigned char>> used_d;
    for (int b = 0; b < ncta; ++b) {
        const int ra = row_start[b], rb = row_start[b + 1], R = rb - ra;
        const int sa = rp[ra], ns = rp[rb] - sa;
        int* tab = jw + (size_t)b * kLzSliceTab;
        std::fill(tab, tab + kLzSliceTab, 0);
        int acc = 0;
        for (int w = 0; 32 * w < R; ++w) {
            const int Lw = jlen[ra + 32 * w];
            tab[2 * w] = acc;
            tab[2 * w + 1] = Lw;
            acc += Lw * kLzSlice;
        }
        positions[b] = acc;
        slots.clear();
        used_d.resize((size_t)R);
        for (int t = 0; t < R; ++t) {
            const int row = jrow[ra + t], s0 = rp[row], len = rp[row + 1] - s0;
            used_d[t].assign((size_t)len, 0);
            for (int d = 0; d < len; ++d) slots.push_back({inv[col[(size_t)s0 + d]], t, eid[(size_t)s0 + d]});
        }
        std::stable_sort(slots.begin(), slots.end(), [](const Slot& x, const Slot& y) { return x.col < y.col; });
        for (int g0 = 0; g0 < ns; g0 += 16) {
            unsigned int banks = 0;
            for (int q = g0; q < std::min(ns, g0 + 16); ++q) {
                const int t = slots[q].t, base = tab[2 * (t >> 5)] + (t & 31);
                std::vector<unsigned char>& u = used_d[t];
                int pick = -1, fallback = -1;
                for (int d = 0; d < (int)u.size(); ++d) {
                    if (u[d]) continue;
                    if (fallback < 0) fallback = d;
                    if (!bankfit) break;
                    if (!((banks >> ((base + d * kLzSlice) & 15)) & 1u)) {
                        pick = d;
                        break;
                    }
                }
                if (pick < 0) pick = fallback;
                u[pick] = 1;
                const int pos = base + pick * kLzSlice;
                banks |= 1u << (pos & 15);
                jcol[(size_t)sa + q] = slots[q].col | (pos << 17);
                jeid[(size_t)sa + q] = slots[q].eid;
            }
        }
    }
}

// Buffers of the on-device Rayleigh-Ritz (lz_rr_main): private copies of T_k, work arrays, result record, stop flag.
void alloc_device_rr(macb_ctx* c) {
    if (c->dev_rr || getenv("MACB_HOST_RR")) return;
    const size_t cap2 = (size_t)c->basis_cap + 4;
    c->d_rr_a = dalloc<double>(cap2);
    c->d_rr_b = dalloc<double>(cap2);
    c->d_rr_b2 = dalloc<double>(cap2);
    c->d_rr_binv = dalloc<double>(cap2);
    c->d_rr_s = dalloc<double>(cap2);
    c->d_rr_w = dalloc<double>(2 * cap2);
    c->d_rr_out = dalloc<RrOut>(1);
    c->d_dev_stop = dalloc<int>(1);
    CK(cudaMallocHost(&c->h_rr, sizeof(RrOut)));
    CK(cudaMemsetAsync(c->d_dev_stop, 0, sizeof(int), c->stream));
    c->dev_rr = true;
}
size_t rr_smem_bytes(const macb_ctx* c) {
    return std::min<size_t>((size_t)220 * 1024, (size_t)64 * ((size_t)c->basis_cap + 32));
}

void setup_persist(macb_ctx* c) {
    const int n = c->n, W = c->W;
    std::vector<int> rs;
    // small graphs: the whole Lanczos state in one SM's shared memory (k_lanczos_small2)
    const size_t small_bytes = (size_t)32 * n + (size_t)16 * c->nnz;
    if (c->persist_v == 3 && small_bytes <= (size_t)224 * 1024 && c->nnz <= (int64_t)kSmallSlots * kPBlock &&
        n <= kSmallRows * kPBlock && !getenv("MACB_NO_SMALL")) {
        c->persist_v = 4;
        c->p_ncta = 1;
        c->slots_smem = small_bytes;
        alloc_device_rr(c);
        if (c->dev_rr) c->slots_smem = std::max(small_bytes, rr_smem_bytes(c));
        raise_dyn_smem((const void*)k_lanczos_small2, c->slots_smem);
        rs.assign(2, n);
        rs[0] = 0;
    }
    if (c->persist_v == 3) {
        // slot-parallel kernel: CTAs sized by work (one slot = 1, one row = 4), chunks of <= kPBlock rows and
        // <= cap slots so that a chunk's products fit in shared memory
        const int64_t cap = (227 * 1024 - 4096) / 8;
        int64_t maxrow = 0;
        for (int i = 0; i < n; ++i) maxrow = std::max<int64_t>(maxrow, c->h_rp[i + 1] - c->h_rp[i]);
        if (maxrow > cap) {
            c->persist_v = 1;   // a single row does not fit the staging buffer: use the row-parallel kernel
        } else {
            const int64_t total = (int64_t)c->nnz + 4 * (int64_t)n;
            // work quantum per CTA: the all-to-all record exchange costs about the same for 20 or 148 CTAs, so small
            // graphs are spread over many SMs (the counter barrier of the older engines preferred few)
            int64_t quantum = 4 * kPBlock;
            if (const char* env = getenv("MACB_CTA_QUANTUM")) quantum = std::max(64, atoi(env));
            c->p_ncta = (int)std::max<int64_t>(1, std::min<int64_t>((total + quantum - 1) / quantum, c->sm_count - 1));   // one SM stays free for the Rayleigh-Ritz CTA
            rs.assign((size_t)c->p_ncta + 1, n);
            rs[0] = 0;
            int row = 0;
            for (int b = 1; b < c->p_ncta; ++b) {
                const int64_t target = total * b / c->p_ncta;
                while (row < n && (int64_t)c->h_rp[row] + 4 * (int64_t)row < target) ++row;
                rs[b] = row;
            }
            std::vector<int> chunk_ptr((size_t)c->p_ncta + 1, 0), chunk_row;
            int64_t max_slots = 1;
            for (int b = 0; b < c->p_ncta; ++b) {
                chunk_ptr[b] = (int)chunk_row.size();
                int r = rs[b];
                while (r < rs[b + 1]) {
                    chunk_row.push_back(r);
                    int e = r;
                    while (e < rs[b + 1] && e - r < kPBlock && (int64_t)c->h_rp[e + 1] - c->h_rp[r] <= cap) ++e;
                    max_slots = std::max<int64_t>(max_slots, (int64_t)c->h_rp[e] - c->h_rp[r]);
                    r = e;
                }
            }
            chunk_ptr[c->p_ncta] = (int)chunk_row.size();
            chunk_row.push_back(n);
            // chunk_row must give, for chunk q, its end as chunk_row[q + 1]: true inside a CTA; at a CTA boundary
            // the next CTA's first chunk starts exactly where this one ends (ranges are contiguous)
            c->d_chunk_ptr = dalloc<int>(chunk_ptr.size());
            c->d_chunk_row = dalloc<int>(chunk_row.size());
            CK(cudaMemcpyAsync(c->d_chunk_ptr, chunk_ptr.data(), sizeof(int) * chunk_ptr.size(), cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(c->d_chunk_row, chunk_row.data(), sizeof(int) * chunk_row.size(), cudaMemcpyHostToDevice, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            c->slots_smem = (size_t)max_slots * sizeof(double);
            c->slots_prod_cap = (int)max_slots;
            // one chunk per CTA and room for 4 more bytes per slot => keep the column indices in shared memory
            bool single = true;
            for (int b = 0; b < c->p_ncta; ++b) single = single && (chunk_ptr[b + 1] - chunk_ptr[b] <= 1);
            if (single && (size_t)max_slots * 12 <= (size_t)224 * 1024 && !getenv("MACB_NO_COLCACHE")) {
                c->slots_cache_cols = 1;
                c->slots_smem = (size_t)max_slots * 12;
            }
            raise_dyn_smem((const void*)k_lanczos_slots, (size_t)(c->slots_smem));
            // sliced layout of k_lanczos_pipe: one chunk per CTA, products + column cache fit, and the last ceil(ncta / 32)
            // warps of every CTA free of rows (they poll the exchange records)
            bool pipe_ok = !getenv("MACB_NO_PIPE") && c->p_ncta <= 256 && n < (1 << 17) - 1;
            for (int b = 0; b < c->p_ncta && pipe_ok; ++b) pipe_ok = rs[b + 1] - rs[b] <= (kPWarps - (c->p_ncta + 31) / 32) * 32;
            const int64_t cap4 = std::max<int64_t>((max_slots + 3) / 4 * 4, 4);
            if (pipe_ok && single && c->slots_cache_cols && !getenv("MACB_NO_JDS") && !c->h_col.empty()) {
                const int ncta = c->p_ncta;
                std::vector<int> jrow((size_t)n), jlen((size_t)n), jcol((size_t)c->nnz), jeid((size_t)c->nnz),
                    jd((size_t)ncta * kLzSliceTab, 0), positions((size_t)ncta, 0);
                build_slice_layout(n, c->h_rp.data(), c->h_col.data(), c->h_eid.data(), ncta, rs.data(), !getenv("MACB_NO_BANKFIT"),
                                   jrow.data(), jlen.data(), jcol.data(), jeid.data(), jd.data(), positions.data());
                const int64_t pos_cap = std::max<int64_t>((*std::max_element(positions.begin(), positions.end()) + 3) / 4 * 4, 4);
                pipe_ok = pos_cap < (1 << 15) && (size_t)pos_cap * 8 + (size_t)cap4 * 4 + kLzSliceTab * 4 <= (size_t)224 * 1024;
                if (pipe_ok) {
                c->d_jrow = dalloc<int>(n);
                c->d_jlen = dalloc<int>(n);
                c->d_jcol = dalloc<int>(c->nnz);
                c->d_jeid = dalloc<int>(c->nnz);
                c->d_jval = dalloc<double>(c->nnz);
                c->d_jd = dalloc<int>(jd.size());
                c->d_xrec = dalloc<double>((size_t)8 * ncta * ncta);
                CK(cudaMemcpyAsync(c->d_jrow, jrow.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
                CK(cudaMemcpyAsync(c->d_jlen, jlen.data(), sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
                CK(cudaMemcpyAsync(c->d_jcol, jcol.data(), sizeof(int) * c->nnz, cudaMemcpyHostToDevice, c->stream));
                CK(cudaMemcpyAsync(c->d_jeid, jeid.data(), sizeof(int) * c->nnz, cudaMemcpyHostToDevice, c->stream));
                CK(cudaMemcpyAsync(c->d_jd, jd.data(), sizeof(int) * jd.size(), cudaMemcpyHostToDevice, c->stream));
                CK(cudaStreamSynchronize(c->stream));
                c->pipe_pos_cap = (int)pos_cap;
                c->pipe_slot_cap = (int)cap4;
                c->pipe_smem = (size_t)pos_cap * 8 + (size_t)cap4 * 4 + kLzSliceTab * 4;
                c->pipe = true;
                {   // gathers in flight per thread: the batch size whose last batch of a step is fullest (see kernels.cuh)
                    const double pt = (double)max_slots / (double)kPBlock;
                    double best = -1.0;
                    for (int vb = 3; vb <= 8; ++vb) {
                        const int nb = std::max(1, (int)std::ceil(pt / vb));
                        const double fill = (pt - (double)(nb - 1) * vb) / vb;
                        if (fill > best + 1e-9 || (std::fabs(fill - best) <= 1e-9 && vb > c->vec_batch)) {
                            best = fill;
                            c->vec_batch = vb;
                        }
                    }
                    if (const char* env = getenv("MACB_VEC_BATCH")) {
                        const int vb = atoi(env);
                        if (vb >= 3 && vb <= 8) c->vec_batch = vb;
                    }
                }
                c->d_zprev = dalloc<double>((size_t)n);
                alloc_device_rr(c);
                // the Rayleigh-Ritz CTA keeps T_k (seven arrays) in the launch's dynamic shared memory while it fits
                if (c->dev_rr) c->pipe_smem = std::max(c->pipe_smem, rr_smem_bytes(c));
                raise_dyn_smem((const void*)k_lanczos_pipe<3>, c->pipe_smem);
                raise_dyn_smem((const void*)k_lanczos_pipe<4>, c->pipe_smem);
                raise_dyn_smem((const void*)k_lanczos_pipe<5>, c->pipe_smem);
                raise_dyn_smem((const void*)k_lanczos_pipe<6>, c->pipe_smem);
                raise_dyn_smem((const void*)k_lanczos_pipe<7>, c->pipe_smem);
                raise_dyn_smem((const void*)k_lanczos_pipe<8>, c->pipe_smem);
                c->persist_v = 5;
                if (!getenv("MACB_NO_L2PIN")) {
                    // keep the weights the Lanczos kernel streams every step (8 bytes per slot) in the persisting part of L2
                    cudaDeviceProp prop;
                    CK(cudaGetDeviceProperties(&prop, c->device));
                    const size_t bytes = std::min<size_t>((size_t)c->nnz * sizeof(double), (size_t)prop.accessPolicyMaxWindowSize);
                    const size_t carve = std::min<size_t>(bytes + (bytes >> 2), (size_t)prop.persistingL2CacheMaxSize);
                    bool carved = false;
                    if (bytes > 0 && carve >= bytes) {
                        std::lock_guard<std::mutex> lk(g_l2_mutex);
                        L2Pin& pin = g_l2_pins[c->device];
                        carved = carve <= pin.carve || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess;
                        if (carved) {
                            pin.carve = std::max(pin.carve, carve);
                            pin.refs++;
                            c->l2_pinned = true;
                        }
                    }
                    if (carved) {
                        cudaStreamAttrValue attr;
                        memset(&attr, 0, sizeof(attr));
                        attr.accessPolicyWindow.base_ptr = c->d_jval;
                        attr.accessPolicyWindow.num_bytes = bytes;
                        attr.accessPolicyWindow.hitRatio = 1.0f;
                        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                        if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
                    } else {
                        cudaGetLastError();
                    }
                }
                if (c->have_x) {   // L(x) was assembled before the engine existed: fill the engine's copy of the weights once
                    k_assemble_jds<<<c->grid_for(c->nnz), kBlock, 0, c->stream>>>(c->nnz, c->d_jeid, c->d_ew, c->d_jval);
                    CK(cudaGetLastError());
                    c->c_launches++;
                }
                }   // pipe_ok (positions fit)
            }
        }
    }
    std::vector<int32_t>().swap(c->h_col);
    std::vector<int32_t>().swap(c->h_eid);
    if (c->persist_v != 3 && c->persist_v != 4 && c->persist_v != 5) {
        // CTAs: one per SM at most (cooperative launch => all co-resident); small graphs use fewer so that the
        // grid barrier stays cheap.
        int64_t want = ((int64_t)c->n * c->W + kPBlock - 1) / kPBlock;
        c->p_ncta = (int)std::max<int64_t>(1, std::min<int64_t>(want, c->sm_count));
        // contiguous row ranges balanced by the number of 4W-slot passes a row needs plus its length
        std::vector<int64_t> cost((size_t)n + 1, 0);
        for (int i = 0; i < n; ++i) {
            int64_t len = c->h_rp[i + 1] - c->h_rp[i];
            int64_t passes = std::max<int64_t>(1, (len + 4 * W - 1) / (4 * W));
            cost[i + 1] = cost[i] + passes * 4 * W + len;
        }
        rs.assign((size_t)c->p_ncta + 1, n);
        rs[0] = 0;
        int row = 0;
        for (int b = 1; b < c->p_ncta; ++b) {
            const int64_t target = cost[n] * b / c->p_ncta;
            while (row < n && cost[row] < target) ++row;
            rs[b] = row;
        }
    }
    c->check_div = (c->nnz >= 1000000) ? 48 : 8;   // steps of >= 10 us leave the host time for frequent checks
    if (const char* env = getenv("MACB_CHECK_DIV")) c->check_div = std::max(1, atoi(env));
    c->d_row_start = dalloc<int>(c->p_ncta + 1);
    CK(cudaMemcpyAsync(c->d_row_start, rs.data(), sizeof(int) * (c->p_ncta + 1), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (c->persist_v == 5) {
        // k_lanczos_pipe: one block of five rows of ld doubles (z buffers, u buffers, diagonal); see the kernel
        c->d_sect[0] = dalloc<double>((size_t)c->ld * 5);
        c->d_sect[1] = c->d_sect[0] + c->ld;
        c->sect_joint = true;
    } else {
        c->d_sect[0] = dalloc<double>((size_t)c->n * 4);
        c->d_sect[1] = dalloc<double>((size_t)c->n * 4);
    }
    c->d_precs = dalloc<LzPartRec>((size_t)c->p_ncta * 2);
    c->d_pst = dalloc<LzPersistState>(1);
    CK(cudaMemsetAsync(c->d_pst, 0, sizeof(LzPersistState), c->stream));
    CK(cudaHostAlloc(&c->h_ab, sizeof(double) * 2 * (c->basis_cap + 2), cudaHostAllocMapped));
    CK(cudaHostAlloc(&c->h_stop, sizeof(int) * 16, cudaHostAllocMapped));
    *c->h_stop = 0;
    for (int64_t j = 0; j < 2 * (c->basis_cap + 2); ++j) c->h_ab[j] = std::numeric_limits<double>::quiet_NaN();
#ifdef MACB_PTIMING
    c->d_ptiming = dalloc<long long>((size_t)64 * c->p_ncta * 9);
    CK(cudaMemsetAsync(c->d_ptiming, 0, sizeof(long long) * 64 * c->p_ncta * 9, c->stream));
#endif
}

void ensure_basis(macb_ctx* c, int max_steps) {
    if (c->d_basis) return;
    // Capacity (Lanczos steps per cycle; a cycle that exhausts it restarts from its Ritz vector): what the caller asked
    // for, never more than n - 1 (the Krylov space on 1-perp is exhausted by then), within a memory budget of
    // MACB_BASIS_GB (default 2 GB, and at most half of the free device memory) -- several handles may share one GPU.
    double gb = 2.0;
    if (const char* env = getenv("MACB_BASIS_GB")) gb = atof(env);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = (size_t)16 << 30; }
    const double budget = std::min(gb * 1073741824.0, 0.5 * (double)free_b);
    const int64_t by_mem = (int64_t)(budget / (8.0 * c->ld)) - 2;
    const int64_t want = std::max<int64_t>(1, std::min<int64_t>(max_steps > 0 ? max_steps : 20000, (int64_t)c->n - 1));
    int64_t cap = std::max<int64_t>(2 * kGraphSteps, std::min<int64_t>(std::min<int64_t>(want, by_mem), 65536));
    cap = ((cap + kGraphSteps - 1) / kGraphSteps) * kGraphSteps;
    c->basis_cap = cap;
    c->d_basis = dalloc<double>((size_t)(cap + 2) * c->ld);   // the kernels write u_{j+1} at the end of phase j
    c->d_alpha = dalloc<double>(cap + 1);
    c->d_beta = dalloc<double>(cap + 2);
    c->d_coef = dalloc<double>(cap + 1);
    CK(cudaMallocHost(&c->h_alpha, sizeof(double) * (cap + 1)));
    CK(cudaMallocHost(&c->h_beta, sizeof(double) * (cap + 2)));
    setup_persist(c);
}

struct FiedlerResult {
    double lambda2 = 0.0, resid = 0.0;
    int steps = 0;
    bool converged = false;
};

// Ritz vector of T_k -> d_v (unit norm, zero mean), true residual test of nx:243.
void finalize_ritz(macb_ctx* c, int k, const std::vector<double>& s, FiedlerResult& out) {
    std::vector<double> coef(k);
    for (int t = 0; t < k; ++t) coef[t] = s[t] / c->h_beta[t];
    CK(cudaMemcpyAsync(c->d_coef, coef.data(), sizeof(double) * k, cudaMemcpyHostToDevice, c->stream));
    k_ritz<<<c->grid_for((int64_t)c->n * 2), kBlock, 0, c->stream>>>(c->n, c->ld, k, c->d_basis, c->d_coef, c->d_v, c->d_sc, c->ws(),
                                                        c->persist_v == 5 ? c->d_jrow : nullptr, nullptr);
    k_center_normalize<<<c->grid_for(c->n), kBlock, 0, c->stream>>>(c->n, c->d_v, c->d_sc);
    launch_spmv<2>(c, c->d_v, c->d_y);
    k_resid_l1<<<c->grid_for(c->n), kBlock, 0, c->stream>>>(c->n, c->d_v, c->d_y, c->d_sc, c->ws());
    CK(cudaGetLastError());
    c->c_launches += 4;
    c->c_spmv += 1;
    CK(cudaMemcpyAsync(c->h_sc, c->d_sc, sizeof(LzScalars), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));  // also keeps `coef` alive until the copy is done
    out.lambda2 = c->h_sc->vLv / c->h_sc->vv;
    out.resid = c->h_sc->res1 / (std::sqrt(c->h_sc->vv) * c->lnorm);
    c->have_prev_v = true;
}


// Check points of the asynchronous Rayleigh-Ritz: a function of k alone, so that the step count at which a solve
// stops -- and with it the result, to the last bit -- does not depend on host/device timing.
// A check costs the host ~0.055 us * k (Sturm multisection + eigenvector of T_k) and runs beside the kernel, whose
// steps take 3.7 us (single-SM kernel) to 12 us (headline size): an interval of k/48 keeps the host below ~25 % duty at
// the headline size, so it never falls behind, and the expected overshoot is k/96 steps (1 %) instead of 6 % at k/8.  That holds
// for graphs that fill the GPU; on pose graphs (steps of 3.7-5.5 us, T_k with lambda_2/lambda_max ~ 1e-5 and thousands
// of steps) the host would be the bottleneck, so they keep k/8 (macb_ctx::check_div, a function of the graph only).
inline int next_check(int k, int div) { return k + std::max(div >= 24 ? 4 : 16, (k / div) & ~3); }

// One Lanczos cycle on the persistent engine with the host Rayleigh-Ritz running concurrently with the kernel.
// Returns: 1 converged (result in `out`, vector in d_v), 0 cycle exhausted without convergence (best Ritz vector
// in d_v), and updates total_steps.
int lanczos_cycle_async(macb_ctx* c, double tol, int k_limit, double brk, int& total_steps, std::vector<double>& s,
                        FiedlerResult& out) {
    const int n = c->n;
    const double sqrtn = std::sqrt((double)n), lnorm = c->lnorm;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    static const bool trace = getenv("MACB_TRACE") != nullptr;
    const auto T0 = std::chrono::steady_clock::now();
    auto us = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - T0).count(); };
    double t_wait = 0.0, t_rr = 0.0;
    int n_checks = 0;
    volatile double* ab = c->h_ab;
    int phases_done = 0;   // phases the device has completed in earlier launches of this cycle
    int k_next = std::min(16, k_limit);
    int k_seen = 0;        // alpha/beta copied into h_alpha/h_beta for indices < k_seen
    double theta_prev = std::numeric_limits<double>::infinity(), theta_delta = -1.0;
    bool invariant = false;
    static const bool adaptive_checks = !getenv("MACB_FIXED_CHECKS");
    double t_launch = 0.0;
    int k_launch = 0;
    double est_prev = 0.0;
    int k_prev = 0;
    while (true) {
        // ---- launch (or resume) the kernel for everything that is left of this cycle
        const int nph = (k_limit + 1) - phases_done;
        if (nph > 0) {
            // entries at and beyond the resume point that an earlier solve / launch may have written
            for (int j = phases_done; j < std::min(c->ab_dirty, k_limit + 1); ++j) {
                ab[2 * j] = nan;
                ab[2 * j + 1] = nan;
            }
            c->ab_dirty = phases_done;
            *(volatile int*)c->h_stop = 0;
            launch_persist(c, nph, true);
            t_launch = us();
            k_launch = phases_done;
        }
        bool stopped = false;
        int k_conv = -1;
        while (!stopped) {
            // wait for beta[k_next] (phase k_next) or for the kernel to end
            const int need = std::min(k_next, k_limit);
            bool kernel_done = false;
            const double tw0 = us();
            // A far-away check point: sleep through half of the predicted wait instead of burning a core (several handles
            // sweeping budgets concurrently share the host's cores with each other's Rayleigh-Ritz).  The step time is
            // measured on this very launch; the stop decision does not depend on when the host looks.
            if (k_seen > 32 && std::isnan(ab[2 * need + 1])) {
                int latest = k_seen;
                while (latest < need && !std::isnan(ab[2 * latest + 1])) ++latest;
                const double us_per_step = (tw0 - t_launch) / (double)std::max(latest - k_launch, 1);
                const double wait_us = 0.5 * us_per_step * (double)(need - latest);
                if (latest > k_launch + 16 && wait_us > 1000.0)
                    std::this_thread::sleep_for(std::chrono::microseconds((long long)std::min(wait_us, 20000.0)));
            }
            // spin on the mapped memory; ask the driver whether the kernel has ended only now and then (the query
            // takes the driver lock: with several handles sweeping budgets concurrently that lock is the bottleneck)
            for (unsigned int spin = 1; std::isnan(ab[2 * need + 1]); ++spin) {
                if ((spin & 1023u) == 0 && cudaStreamQuery(c->stream) != cudaErrorNotReady) {
                    kernel_done = true;
                    break;
                }
                __builtin_ia32_pause();
            }
            // Every exit from here on that is not a normal return stops the kernel, waits for it and records how far it
            // wrote into the mapped coefficient array (the next solve re-poisons exactly that range).
            auto abort_cycle = [&](const char* what) {
                *(volatile int*)c->h_stop = 1;
                cudaStreamSynchronize(c->stream);
                c->ab_dirty = std::max(c->ab_dirty, std::max((int)((volatile int*)c->h_stop)[1], k_limit + 1));
                throw ArgFail{what, MACB_ERR_STATE};
            };
            if (kernel_done && std::isnan(ab[2 * need + 1])) {
                const cudaError_t se = cudaStreamSynchronize(c->stream);  // surfaces launch/runtime errors
                if (se != cudaSuccess) {
                    c->ab_dirty = std::max(c->ab_dirty, k_limit + 1);
                    throw CudaFail{se, "cudaStreamSynchronize (Lanczos kernel)", __LINE__};
                }
                if (std::isnan(ab[2 * need + 1])) abort_cycle("macb_fiedler: Lanczos kernel ended early");
            }
            const double tw1 = us();
            t_wait += tw1 - tw0;
            // The kernel publishes the entries with plain stores to mapped memory, one phase after the other; nothing orders
            // them with respect to each other on the way to the host, so beta[need] being visible does not make the earlier
            // entries visible: wait for every entry on its own (NaN = not there yet; a poisoned recurrence publishes +inf).
            for (int j = k_seen; j <= need; ++j) {
                for (unsigned int spin = 1; std::isnan(ab[2 * j]) || std::isnan(ab[2 * j + 1]); ++spin) {
                    if ((spin & 0xfffffu) == 0 && cudaStreamQuery(c->stream) != cudaErrorNotReady &&
                        (std::isnan(ab[2 * j]) || std::isnan(ab[2 * j + 1])))
                        abort_cycle("macb_fiedler: a Lanczos coefficient never reached the host");
                    __builtin_ia32_pause();
                }
                c->h_alpha[j] = ab[2 * j];
                c->h_beta[j] = ab[2 * j + 1];
                if (!std::isfinite(c->h_alpha[j]) || !std::isfinite(c->h_beta[j]))
                    abort_cycle("macb_fiedler: the Lanczos recurrence produced a non-finite coefficient");
            }
            k_seen = std::max(k_seen, need + 1);
            int k = need;
            for (int j = 1; j <= need; ++j)
                if (!(c->h_beta[j] > brk)) {
                    k = j;
                    invariant = true;
                    break;
                }
            if (k == 0) abort_cycle("macb_fiedler: Lanczos made no progress");
            const double theta = tridiag_smallest_value(c->h_alpha, c->h_beta, k,
                                                        invariant ? std::numeric_limits<double>::infinity() : theta_prev, theta_delta);
            s.resize(k);
            tridiag_vector(c->h_alpha, c->h_beta, k, theta, s.data());
            if (std::isfinite(theta_prev)) theta_delta = 2.0 * std::fabs(theta_prev - theta);
            theta_prev = theta;
            const double est = std::fabs(c->h_beta[k]) * std::fabs(s[k - 1]);
            const bool exhausted = invariant || need >= k_limit;
            t_rr += us() - tw1;
            ++n_checks;
            // ||r||_1 <= sqrt(n) ||r||_2 holds with equality only for a flat residual; the residual of a Ritz pair is a
            // multiple of the next Lanczos vector, whose entries are Gaussian-like: ||r||_1 ~ 0.80 sqrt(n) ||r||_2.  Ask
            // for 10 % margin on that prediction; the true residual is tested below in any case.
            if (0.88 * est * sqrtn < tol * lnorm || exhausted) {
                *(volatile int*)c->h_stop = 1;
                const double ts0 = us();
                CK(cudaStreamSynchronize(c->stream));
                if (trace)
                    fprintf(stderr, "[macb] k=%d checks=%d wait=%.0fus rr=%.0fus stop->sync=%.0fus t=%.0fus\n", k, n_checks, t_wait, t_rr,
                            us() - ts0, us());
                const int phases_before = phases_done;
                phases_done = ((volatile int*)c->h_stop)[1];   // written by the kernel at exit (host-mapped)
                c->ab_dirty = std::max(c->ab_dirty, phases_done);
                if (c->bench_time_iters) {
                    float ms = 0.f;
                    CK(cudaEventElapsedTime(&ms, c->lz0, c->lz1));
                    c->lz_kernel_ms += ms;
                    c->lz_kernel_phases += phases_done - phases_before;
                }
                stopped = true;
                k_conv = k;
                finalize_ritz(c, k, s, out);
                if (trace) fprintf(stderr, "[macb] phases=%d finalize done t=%.0fus resid=%.2e\n", phases_done, us(), out.resid);
                if (out.resid < tol) {
                    total_steps += phases_done;
                    c->c_steps += phases_done;
                    c->c_spmv += phases_done;
                    out.converged = true;
                    return 1;
                }
                if (exhausted || phases_done >= k_limit + 1) {
                    total_steps += phases_done;
                    c->c_steps += phases_done;
                    c->c_spmv += phases_done;
                    return invariant ? -1 : 0;
                }
                // estimate passed but the true residual did not: resume the kernel and look again a little later
                k_next = next_check(need, c->check_div);
            } else {
                // Next check: the residual estimate of a resolved Ritz pair decays geometrically, so two consecutive
                // estimates predict where it crosses the threshold; look again half-way there (Lanczos converges
                // superlinearly, the prediction errs on the late side), never closer than 4 steps, never further than
                // the fixed k/div schedule allows while the estimate is not yet decreasing.  The schedule depends on
                // (alpha, beta) only -- not on timing -- so a solve stays a pure function of its input.
                const double target = tol * lnorm / (0.88 * sqrtn);
                int nk = next_check(need, c->check_div);
                if (adaptive_checks && est_prev > 0.0 && est > target && est < est_prev && k > k_prev) {
                    const double slope = (std::log(est) - std::log(est_prev)) / (double)(k - k_prev);
                    const double pred = (std::log(target) - std::log(est)) / slope;   // steps still to go at this rate
                    const double cap = std::max(16.0, 0.25 * (double)k);
                    nk = need + (int)std::max(4.0, std::min(0.5 * pred, cap));
                }
                k_prev = k;
                est_prev = est;
                k_next = nk;
            }
        }
        (void)k_conv;
    }
}

// ---- eigen-solve with the stop decision on the device (k_lanczos_pipe + its Rayleigh-Ritz CTA) -------------------------------
// Everything is enqueued on the handle's stream and NOTHING is read back: start vector -> z_0 = L u_0 -> ONE cooperative launch
// (solver CTAs + the Rayleigh-Ritz CTA that stops them) -> Ritz vector with the order k and the coefficients the device left
// -> normalisation -> Rayleigh quotient and the reference's residual (nx:243).  The caller synchronises when IT needs the
// numbers (macb_fw_run: once per Frank-Wolfe iteration) and then calls finish_fiedler_device.
bool device_fiedler_available(const macb_ctx* c) {
    return c->dev_rr && (c->persist_v == 5 || c->persist_v == 4);
}

void enqueue_fiedler_device(macb_ctx* c, double tol, int max_steps, bool use_warm) {
    const int n = c->n;
    const int k_lim = (int)std::min<int64_t>(c->basis_cap, (int64_t)std::min(max_steps, n - 1));
    const double* src = (use_warm && c->have_prev_v) ? c->d_v : c->d_x0;
    const bool small = c->persist_v == 4;
    if (small) {
        k_lz_persist_init<<<c->grid_for(n), kBlock, 0, c->stream>>>(n, src, c->d_diag, c->d_sect[0], c->d_pst, c->d_precs, 2 * c->p_ncta,
                                                                     nullptr, nullptr, 0);
        k_rr_reset<<<c->grid_for(k_lim + 2), kBlock, 0, c->stream>>>(c->d_alpha, c->d_beta, k_lim + 2, c->d_rr_out, c->d_dev_stop);
    } else {
        launch_spmv<0>(c, src, c->d_y);   // z_0 = L u_0 (the shift is applied by the init kernel)
        k_lz_pipe_init<<<c->grid_for(n), kBlock, 0, c->stream>>>(n, src, c->d_y, c->d_jrow, c->d_sc, c->d_sect[0], c->d_sect[1], c->d_basis,
                                                                  c->d_xrec, (int64_t)8 * c->p_ncta * c->p_ncta, c->d_pst, c->d_alpha,
                                                                  c->d_beta, k_lim + 2, c->d_rr_out, c->d_dev_stop);
    }
    CK(cudaGetLastError());
    RrArgs R;
    R.alpha = c->d_alpha; R.beta = c->d_beta;
    R.a = c->d_rr_a; R.b = c->d_rr_b; R.b2 = c->d_rr_b2; R.binv = c->d_rr_binv; R.s = c->d_rr_s;
    {
        const size_t cap2 = (size_t)c->basis_cap + 4;
        R.dp = c->d_rr_w; R.dm = c->d_rr_w + cap2;
    }
    R.coef = c->d_coef; R.out = c->d_rr_out; R.dev_stop = c->d_dev_stop; R.sc = c->d_sc;
    R.tol = tol; R.n = n; R.k_limit = k_lim; R.check_div = c->check_div; R.enabled = 1;
    c->rr_launch = R;
    launch_persist(c, k_lim + 1, true);
    c->rr_launch.enabled = 0;
    k_ritz<<<c->grid_for((int64_t)n * 2), kBlock, 0, c->stream>>>(n, c->ld, 0, c->d_basis, c->d_coef, c->d_v, c->d_sc, c->ws(), small ? nullptr : c->d_jrow,
                                                        c->d_rr_out);
    k_center_normalize<<<c->grid_for(n), kBlock, 0, c->stream>>>(n, c->d_v, c->d_sc);
    launch_spmv<2>(c, c->d_v, c->d_y);
    k_resid_l1<<<c->grid_for(n), kBlock, 0, c->stream>>>(n, c->d_v, c->d_y, c->d_sc, c->ws());
    CK(cudaGetLastError());
    c->c_launches += 7;
    c->c_spmv += 2;
    c->c_solves++;
    c->have_v = true;
    c->have_prev_v = true;
}

// After the stream has been synchronised with h_sc / h_rr copied back: the numbers of the solve.  false = the device-side
// decision did not produce a pair that passes the reference's residual test (the estimate was optimistic, the cycle ran out
// of basis, the eigenvector recurrence failed, ...): the caller repeats the solve on the host-driven path.
bool finish_fiedler_device(macb_ctx* c, double tol, FiedlerResult& out) {
    const RrOut& r = *c->h_rr;
    c->lnorm = c->h_sc->lnorm;
    c->nnz_active = c->h_sc->nnz_active;
    c->c_steps += r.phases;
    c->c_spmv += r.phases;
    if (c->bench_time_iters) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, c->lz0, c->lz1));
        c->lz_kernel_ms += ms;
        c->lz_kernel_phases += r.phases;
        c->lz_algo_bytes += (double)r.phases * ((double)(c->nnz_active + c->n) * 12.0 + ((double)c->n + 1.0) * 4.0 + 40.0 * (double)c->n);
    }
    out.steps = r.phases;
    if (!(c->lnorm > 0.0) || r.status <= 0 || r.k <= 0) return false;
    out.lambda2 = c->h_sc->vLv / c->h_sc->vv;
    out.resid = c->h_sc->res1 / (std::sqrt(c->h_sc->vv) * c->lnorm);
    out.converged = out.resid < tol;
    return out.converged;
}

// Deflated Lanczos on P L(x) P.  Replaces nx:149-253 (see macb200.h).
int run_fiedler(macb_ctx* c, double tol, int max_steps, int warm, FiedlerResult& out) {
    if (!c->have_x) throw ArgFail{"macb_fiedler: call macb_set_x first", MACB_ERR_STATE};
    if (c->n < 2) throw ArgFail{"macb_fiedler: need at least 2 nodes", MACB_ERR_ARG};
    if (max_steps <= 0) max_steps = 20000;
    PhaseTimer pt(c, MACB_T_FIEDLER);
    ensure_basis(c, max_steps);
    const int n = c->n;
    const double sqrtn = std::sqrt((double)n);
    const double lnorm = c->lnorm;
    if (lnorm > 0.0 && device_fiedler_available(c) && !c->force_host_rr) {
        // stop decision on the device: enqueue everything, ONE synchronisation, read the numbers
        enqueue_fiedler_device(c, tol, max_steps, warm != 0);
        CK(cudaMemcpyAsync(c->h_sc, c->d_sc, sizeof(LzScalars), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(c->h_rr, c->d_rr_out, sizeof(RrOut), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        if (finish_fiedler_device(c, tol, out)) return MACB_OK;
        c->c_dev_fallbacks++;   // rare: repeat on the host-driven path below (restarts, twisted factorisation, resume)
        c->c_solves--;
    }
    c->c_solves++;
    if (!(lnorm > 0.0)) {  // empty graph: every vector orthogonal to 1 is a null vector
        CK(cudaMemcpyAsync(c->d_v, c->d_x0, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
        out = FiedlerResult{};
        out.converged = true;
        c->have_v = true;
        return MACB_OK;
    }
    // beta_j below this => span(u_0..u_{j-1}) is invariant to within the requested tolerance: every Ritz
    // pair of T_j then has residual <= beta_j < tol ||L||_inf / sqrt(n), so T_j is final.  (Dividing by a
    // beta that small would also amplify rounding noise into the next vector.)
    const double brk = std::max(1e-12 * lnorm, 0.25 * tol * lnorm / sqrtn);

    bool use_warm = warm && c->have_prev_v;
    int total_steps = 0;
    std::vector<double> s;
    for (int restart = 0; restart < 64; ++restart) {
        // ---- (re)start
        const double* src = use_warm ? c->d_v : c->d_x0;
        if (c->persist_v == 5) {
            launch_spmv<0>(c, src, c->d_y);   // z_0 = L u_0 (the shift is applied by the init kernel)
            c->c_launches++;
            c->c_spmv++;
            k_lz_pipe_init<<<c->grid_for(n), kBlock, 0, c->stream>>>(n, src, c->d_y, c->d_jrow, c->d_sc, c->d_sect[0], c->d_sect[1],
                                                                      c->d_basis, c->d_xrec, (int64_t)8 * c->p_ncta * c->p_ncta, c->d_pst,
                                                                      c->d_alpha, c->d_beta, 0, nullptr, nullptr);
        } else {
            k_lz_persist_init<<<c->grid_for(n), kBlock, 0, c->stream>>>(n, src, c->d_diag, c->d_sect[0], c->d_pst, c->d_precs,
                                                                         2 * c->p_ncta, nullptr, nullptr, 0);
        }
        CK(cudaGetLastError());
        c->c_launches++;
        {
            const int k_lim = (int)std::min<int64_t>(c->basis_cap, (int64_t)std::min(max_steps - total_steps, n - 1));
            const int rc = lanczos_cycle_async(c, tol, k_lim, brk, total_steps, s, out);
            out.steps = total_steps;
            c->have_v = true;
            if (rc == 1) return MACB_OK;
            use_warm = true;   // explicit restart from the current Ritz vector
            if (rc < 0 || total_steps >= max_steps) break;
        }
    }
    out.steps = total_steps;
    return MACB_NOT_CONVERGED;
}

void launch_gradient(macb_ctx* c) {
    if (!c->have_v) throw ArgFail{"macb_gradient: call macb_fiedler first", MACB_ERR_STATE};
    PhaseTimer pt(c, MACB_T_GRADIENT);
    if (c->m > 0) {
        k_gradient<<<c->grid_for(c->m), kBlock, 0, c->stream>>>(c->m, c->d_ci, c->d_cj, c->d_kappa, c->d_v, c->d_x, c->d_g,
                                                                c->d_sc, c->ws());
        CK(cudaGetLastError());
        c->c_launches++;
    }
    c->have_g = true;
    c->have_sel = false;
}

// radix-select passes over `g` (device, length m) + selection mask into `sel`; dual term into sc
// defer = true: the "are there ties at the k-th value" decision (which needs the selection state on the host) is not
// taken here: the tie-free apply kernel is enqueued speculatively and the caller, after its own synchronisation, calls
// topk_fixup(), which re-runs the ranked path in the rare case that ties straddle the budget.
void launch_topk(macb_ctx* c, const double* g, const double* x, int64_t k, uint8_t* sel, SelState* st = nullptr, bool defer = false) {
    if (!st) st = c->d_sel_state;
    PhaseTimer pt(c, MACB_T_TOPK);
    const int64_t m = c->m;
    if (m == 0) {   // nothing to select from: no selection, dual term 0 (the header allows m = 0 handles)
        k_sel_init<<<1, kBlock, 0, c->stream>>>(st, 0ll);
        k_clear_lp_scalars<<<1, 1, 0, c->stream>>>(c->d_sc);
        c->c_launches += 2;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(c->h_sel_state, st, sizeof(SelState), cudaMemcpyDeviceToHost, c->stream));
        if (!defer) CK(cudaStreamSynchronize(c->stream));
        return;
    }
    const int grid = c->grid_for(m);
    if (k <= 0) {   // nothing to find: "take none" state (for k > 0 the select kernels write every field of the state themselves)
        k_sel_init<<<1, kBlock, 0, c->stream>>>(st, (long long)k);
        c->c_launches++;
    }
    if (k > 0 && m <= kSel2SmallMax) {
        k_sel2_small<<<1, kSel2Block, kSel2Bins * sizeof(unsigned int), c->stream>>>(m, g, (long long)k, st);
        c->c_launches += 1;
    } else if (k > 0) {
        const int grid2 = (int)std::min<int64_t>(c->sm_count, (m + 4095) / 4096);
        k_sel2_hist<<<grid2, kSel2Block, kSel2Bins * sizeof(unsigned int), c->stream>>>(m, g, (long long)k, c->d_sel2_hist, c->d_sel2);
        k_sel2_compact<<<grid, kBlock, 0, c->stream>>>(m, g, c->d_sel2, c->d_sel_cand);
        k_sel2_refine<<<1, kSel2Block, 0, c->stream>>>(c->d_sel2, c->d_sel_cand, st);
        c->c_launches += 3;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(c->h_sel_state, st, sizeof(SelState), cudaMemcpyDeviceToHost, c->stream));
    int64_t chunk = (m + grid - 1) / grid;
    chunk = ((chunk + kBlock - 1) / kBlock) * kBlock;
    const int nblocks = (int)((m + chunk - 1) / chunk);
    if (defer) {
        k_sel_apply<false><<<nblocks, kBlock, 0, c->stream>>>(m, chunk, g, x, st, c->d_blockcnt, sel, c->d_sc, c->ws());
        c->c_launches += 1;
        CK(cudaGetLastError());
        return;
    }
    CK(cudaStreamSynchronize(c->stream));
    const bool ranked = k > 0 && c->h_sel_state->eq_total != c->h_sel_state->remaining;
    if (ranked) {
        k_sel_tie_count<<<nblocks, kBlock, 0, c->stream>>>(m, chunk, g, st, c->d_blockcnt);
        k_sel_tie_scan<<<1, 32, 0, c->stream>>>(nblocks, c->d_blockcnt);
        k_sel_apply<true><<<nblocks, kBlock, 0, c->stream>>>(m, chunk, g, x, st, c->d_blockcnt, sel, c->d_sc, c->ws());
        c->c_launches += 3;
    } else {
        k_sel_apply<false><<<nblocks, kBlock, 0, c->stream>>>(m, chunk, g, x, st, c->d_blockcnt, sel, c->d_sc, c->ws());
        c->c_launches += 1;
    }
    CK(cudaGetLastError());
}

// After a deferred launch_topk and a stream synchronisation: true (and the ranked selection re-done, synchronised) when
// ties at the k-th value straddle the budget.
bool topk_fixup(macb_ctx* c, const double* g, const double* x, int64_t k, uint8_t* sel, SelState* st = nullptr) {
    if (!st) st = c->d_sel_state;
    if (c->m == 0 || !(k > 0 && c->h_sel_state->eq_total != c->h_sel_state->remaining)) return false;
    const int64_t m = c->m;
    const int grid = c->grid_for(m);
    int64_t chunk = (m + grid - 1) / grid;
    chunk = ((chunk + kBlock - 1) / kBlock) * kBlock;
    const int nblocks = (int)((m + chunk - 1) / chunk);
    k_sel_tie_count<<<nblocks, kBlock, 0, c->stream>>>(m, chunk, g, st, c->d_blockcnt);
    k_sel_tie_scan<<<1, 32, 0, c->stream>>>(nblocks, c->d_blockcnt);
    k_sel_apply<true><<<nblocks, kBlock, 0, c->stream>>>(m, chunk, g, x, st, c->d_blockcnt, sel, c->d_sc, c->ws());
    c->c_launches += 3;
    CK(cudaGetLastError());
    return true;
}

void set_x_device(macb_ctx* c, double tol) {
    // d_x already holds x
    c->min_sel_tol = tol;
    if (c->m > 0) {
        k_edge_weights<<<c->grid_for(c->m), kBlock, 0, c->stream>>>(c->m, c->d_x, c->d_kappa, tol, c->d_ew + c->nf);
        CK(cudaGetLastError());
        c->c_launches++;
    }
    launch_assemble(c);
    c->have_x = true;
}

template <typename F>
int guarded(macb_ctx* c, F&& f) {
    if (!c) return MACB_ERR_ARG;
    try {
        CK(cudaSetDevice(c->device));
        return f();
    } catch (const CudaFail& e) {
        char buf[512];
        snprintf(buf, sizeof(buf), "CUDA error %d (%s) at api.cu:%d: %s", (int)e.e, cudaGetErrorString(e.e), e.line, e.what);
        c->err = buf;
        cudaGetLastError();
        return MACB_ERR_CUDA;
    } catch (const ArgFail& e) {
        c->err = e.msg;
        return e.code;
    } catch (const std::bad_alloc&) {
        c->err = "host allocation failed";
        return MACB_ERR_NOMEM;
    }
}

}  // namespace

// ================================================================================================ C-ABI
extern "C" {

const char* macb_version(void) { return "macb200 0.1.0 (sm_100a)"; }

const char* macb_last_error(macb_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int macb_host_build_pattern(int32_t n, int64_t nf, const int32_t* fi, const int32_t* fj, int64_t m, const int32_t* ci,
                            const int32_t* cj, int32_t* row_ptr, int32_t* col, int32_t* eid, int64_t* nnz) {
    try {
        std::vector<int32_t> rp, cc, ee;
        build_pattern(n, nf, fi, fj, m, ci, cj, rp, cc, ee);
        if (row_ptr) std::copy(rp.begin(), rp.end(), row_ptr);
        if (col) std::copy(cc.begin(), cc.end(), col);
        if (eid) std::copy(ee.begin(), ee.end(), eid);
        if (nnz) *nnz = (int64_t)cc.size();
        return MACB_OK;
    } catch (const ArgFail& e) {
        g_create_error = e.msg;
        return e.code;
    }
}

int macb_host_build_slices(int32_t n, const int32_t* rp, const int32_t* col, const int32_t* eid, int32_t ncta, const int32_t* row_start,
                           int bankfit, int32_t* jrow, int32_t* jlen, int32_t* jcol, int32_t* jeid, int32_t* jw, int32_t* positions) {
    if (n < 0 || ncta < 1 || !rp || !row_start || !jrow || !jlen || !jcol || !jeid || !jw || !positions) return MACB_ERR_ARG;
    if (n >= (1 << 17) - 1) return MACB_ERR_ARG;
    for (int b = 0; b < ncta; ++b) {
        const int R = row_start[b + 1] - row_start[b];
        if (R < 0 || R > 32 * (kLzSliceTab / 2)) return MACB_ERR_ARG;
        int64_t pos = 0;
        for (int r = row_start[b]; r < row_start[b + 1]; ++r) pos = std::max<int64_t>(pos, rp[r + 1] - rp[r]);
        if (pos * kLzSlice * ((R + 31) / 32) >= (1 << 15)) {   // cheap upper bound first, exact count below
            std::vector<int> lens;
            for (int r = row_start[b]; r < row_start[b + 1]; ++r) lens.push_back(rp[r + 1] - rp[r]);
            std::sort(lens.begin(), lens.end(), std::greater<int>());
            int64_t exact = 0;
            for (size_t t = 0; t < lens.size(); t += 32) exact += (int64_t)lens[t] * kLzSlice;
            if (exact >= (1 << 15)) return MACB_ERR_ARG;
        }
    }
    try {
        build_slice_layout(n, rp, col, eid, ncta, row_start, bankfit != 0, jrow, jlen, jcol, jeid, jw, positions);
    } catch (const std::bad_alloc&) {
        return MACB_ERR_NOMEM;
    }
    return MACB_OK;
}

int macb_tridiag_smallest(const double* a, const double* b, int k, double* theta, double* s) {
    if (!a || !b || k <= 0) return MACB_ERR_ARG;
    double th = tridiag_smallest_value(a, b, k, std::numeric_limits<double>::infinity());
    if (theta) *theta = th;
    if (s) tridiag_vector(a, b, k, th, s);
    return MACB_OK;
}

int macb_create(int32_t n, int64_t nf, const int32_t* fi, const int32_t* fj, const double* fw, int64_t m,
                const int32_t* ci, const int32_t* cj, const double* ckappa, int device, macb_handle* out) {
    if (!out) return MACB_ERR_ARG;
    *out = nullptr;
    if (n < 1 || nf < 0 || m < 0 || nf + m >= (int64_t)std::numeric_limits<int32_t>::max()) {
        g_create_error = "macb_create: bad sizes";
        return MACB_ERR_ARG;
    }
    macb_ctx* c = new macb_ctx();
    try {
        if (device < 0) CK(cudaGetDevice(&device));
        c->device = device;
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        c->sm_count = prop.multiProcessorCount;
        c->grid_max = c->sm_count * (2048 / kBlock);
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&c->ev0));
        CK(cudaEventCreate(&c->ev1));
        CK(cudaEventCreate(&c->it0));
        CK(cudaEventCreate(&c->it1));
        CK(cudaEventCreate(&c->lz0));
        CK(cudaEventCreate(&c->lz1));
        c->n = n;
        c->ld = ((n + 31) / 32) * 32;
        c->nf = nf;
        c->m = m;

        std::vector<int32_t> rp, col, eid;
        build_pattern(n, nf, fi, fj, m, ci, cj, rp, col, eid);
        c->nnz = (int64_t)col.size();
        c->W = pick_width((double)c->nnz / std::max(1, n));
        c->h_rp = rp;
        c->h_col = col;
        c->h_eid = eid;
        if (const char* env = getenv("MACB_PERSIST_V")) c->persist_v = atoi(env);

        c->d_rp = dalloc<int>(n + 1);
        c->d_col = dalloc<int>(c->nnz);
        c->d_eid = dalloc<int>(c->nnz);
        c->d_val = dalloc<double>(c->nnz);
        c->d_diag = dalloc<double>(n);
        c->d_ew = dalloc<double>(nf + m);
        c->d_ci = dalloc<int>(m);
        c->d_cj = dalloc<int>(m);
        c->d_kappa = dalloc<double>(m);
        c->d_x = dalloc<double>(m);
        c->d_g = dalloc<double>(m);
        c->d_tmp_m = dalloc<double>(m);
        c->d_sel = dalloc<uint8_t>(m);
        c->d_v = dalloc<double>(c->ld);
        c->d_y = dalloc<double>(c->ld);
        c->d_x0 = dalloc<double>(c->ld);
        c->d_tmp_n = dalloc<double>(c->ld);
        c->d_sc = dalloc<LzScalars>(1);
        c->d_partials = dalloc<double>((size_t)c->grid_max * 8);
        c->d_counter = dalloc<unsigned int>(8);
        c->d_sel_state = dalloc<SelState>(1);
        c->d_blockcnt = dalloc<unsigned int>(c->grid_max + 8);
        c->d_sel2_hist = dalloc<unsigned int>(kSel2Bins);
        c->d_sel2 = dalloc<Sel2State>(1);
        c->d_sel_cand = dalloc<unsigned long long>(m);
        CK(cudaMemsetAsync(c->d_sel2_hist, 0, kSel2Bins * sizeof(unsigned int), c->stream));
        CK(cudaMemsetAsync(c->d_sel2, 0, sizeof(Sel2State), c->stream));
        raise_dyn_smem((const void*)k_sel2_hist, (size_t)(kSel2Bins * sizeof(unsigned int)));
        raise_dyn_smem((const void*)k_sel2_small, (size_t)(kSel2Bins * sizeof(unsigned int)));
        CK(cudaMallocHost(&c->h_sc, sizeof(LzScalars)));
        CK(cudaMallocHost(&c->h_sel_state, sizeof(SelState)));
        CK(cudaMemsetAsync(c->d_counter, 0, 8 * sizeof(unsigned int), c->stream));
        CK(cudaMemsetAsync(c->d_sc, 0, sizeof(LzScalars), c->stream));
        CK(cudaMemsetAsync(c->d_x, 0, sizeof(double) * std::max<int64_t>(m, 1), c->stream));
        CK(cudaMemsetAsync(c->d_ew, 0, sizeof(double) * std::max<int64_t>(nf + m, 1), c->stream));

        CK(cudaMemcpyAsync(c->d_rp, rp.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice, c->stream));
        if (c->nnz) {
            CK(cudaMemcpyAsync(c->d_col, col.data(), sizeof(int) * c->nnz, cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(c->d_eid, eid.data(), sizeof(int) * c->nnz, cudaMemcpyHostToDevice, c->stream));
        }
        if (nf) CK(cudaMemcpyAsync(c->d_ew, fw, sizeof(double) * nf, cudaMemcpyHostToDevice, c->stream));
        if (m) {
            CK(cudaMemcpyAsync(c->d_ci, ci, sizeof(int) * m, cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(c->d_cj, cj, sizeof(int) * m, cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(c->d_kappa, ckappa, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
        }
        CK(cudaStreamSynchronize(c->stream));
        upload_start(c, nullptr);
    } catch (const CudaFail& e) {
        char buf[512];
        snprintf(buf, sizeof(buf), "macb_create: CUDA error %d (%s) at api.cu:%d: %s", (int)e.e, cudaGetErrorString(e.e),
                 e.line, e.what);
        g_create_error = buf;
        cudaGetLastError();
        free_all(c);
        return MACB_ERR_CUDA;
    } catch (const ArgFail& e) {
        g_create_error = "macb_create: " + e.msg;
        free_all(c);
        return e.code;
    } catch (const std::bad_alloc&) {
        g_create_error = "macb_create: host allocation failed";
        free_all(c);
        return MACB_ERR_NOMEM;
    }
    *out = c;
    return MACB_OK;
}

int macb_destroy(macb_handle h) {
    if (!h) return MACB_OK;
    free_all(h);
    return MACB_OK;
}

int macb_set_x(macb_handle h, const double* x, double min_sel_tol) {
    return guarded(h, [&]() {
        if (!x && h->m > 0) throw ArgFail{"macb_set_x: x is NULL", MACB_ERR_ARG};
        {
            PhaseTimer pt(h, MACB_T_COPY);
            if (h->m) CK(cudaMemcpyAsync(h->d_x, x, sizeof(double) * h->m, cudaMemcpyHostToDevice, h->stream));
        }
        set_x_device(h, min_sel_tol);
        return (int)MACB_OK;
    });
}

int macb_get_x(macb_handle h, double* x) {
    return guarded(h, [&]() {
        if (h->m) CK(cudaMemcpyAsync(x, h->d_x, sizeof(double) * h->m, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return (int)MACB_OK;
    });
}

int macb_spmv(macb_handle h, const double* v, double* y) {
    return guarded(h, [&]() {
        if (!h->have_x) throw ArgFail{"macb_spmv: call macb_set_x first", MACB_ERR_STATE};
        if (!v || !y) throw ArgFail{"macb_spmv: NULL vector", MACB_ERR_ARG};
        CK(cudaMemcpyAsync(h->d_tmp_n, v, sizeof(double) * h->n, cudaMemcpyHostToDevice, h->stream));
        launch_spmv<0>(h, h->d_tmp_n, h->d_y);
        h->c_launches++;
        h->c_spmv++;
        CK(cudaMemcpyAsync(y, h->d_y, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return (int)MACB_OK;
    });
}

int macb_lnorm(macb_handle h, double* lnorm) {
    return guarded(h, [&]() {
        if (!h->have_x) throw ArgFail{"macb_lnorm: call macb_set_x first", MACB_ERR_STATE};
        if (lnorm) *lnorm = h->lnorm;
        return (int)MACB_OK;
    });
}

int macb_set_start(macb_handle h, const double* x0) {
    return guarded(h, [&]() {
        upload_start(h, x0);
        return (int)MACB_OK;
    });
}

int macb_fiedler(macb_handle h, double tol, int max_steps, int warm, double* lambda2, double* v, int* steps,
                 double* resid) {
    return guarded(h, [&]() {
        FiedlerResult r;
        int rc = run_fiedler(h, tol, max_steps, warm, r);
        if (lambda2) *lambda2 = r.lambda2;
        if (steps) *steps = r.steps;
        if (resid) *resid = r.resid;
        if (v) {
            PhaseTimer pt(h, MACB_T_COPY);
            CK(cudaMemcpyAsync(v, h->d_v, sizeof(double) * h->n, cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
        }
        if (rc == MACB_NOT_CONVERGED) h->err = "macb_fiedler: eigen-iteration did not reach tol within max_steps";
        return rc;
    });
}

// lambda2 of L(x_b) for a batch of iterates, with ONE host synchronisation: every solve (assemble, eigen-solve with the
// stop decision on the device, residual) is enqueued behind the previous one and leaves its scalars in its own pinned slot.
int macb_evaluate_batch(macb_handle h, const double* xs, int nb, double tol, double min_sel_tol, int max_steps, double* lambda2,
                        double* resid) {
    return guarded(h, [&]() {
        if (nb < 0 || (nb > 0 && ((!xs && h->m > 0) || !lambda2))) throw ArgFail{"macb_evaluate_batch: bad arguments", MACB_ERR_ARG};
        if (h->n < 2) throw ArgFail{"macb_evaluate_batch: need at least 2 nodes", MACB_ERR_ARG};
        if (max_steps <= 0) max_steps = 20000;
        int status = MACB_OK;
        ensure_basis(h, max_steps);
        const int64_t m = h->m;
        auto one_by_one = [&](int b) {   // host-driven path (engines without the device-side decision, or a failed decision)
            int rc = macb_set_x(h, xs + (size_t)b * m, min_sel_tol);
            if (rc < 0) throw ArgFail{h->err, rc};
            FiedlerResult fr;
            rc = run_fiedler(h, tol, max_steps, 0, fr);
            if (rc == MACB_NOT_CONVERGED) status = rc;
            lambda2[b] = fr.lambda2;
            if (resid) resid[b] = fr.resid;
        };
        if (!device_fiedler_available(h) || h->profile) {
            for (int b = 0; b < nb; ++b) one_by_one(b);
            return status;
        }
        LzScalars* hs = nullptr;
        RrOut* hr = nullptr;
        double* hx = nullptr;   // pinned staging of the iterates: a pageable source would serialise the copies with the host
        CK(cudaMallocHost(&hs, sizeof(LzScalars) * std::max(nb, 1)));
        CK(cudaMallocHost(&hr, sizeof(RrOut) * std::max(nb, 1)));
        CK(cudaMallocHost(&hx, sizeof(double) * std::max<int64_t>(m, 1) * 2));
        h->min_sel_tol = min_sel_tol;
        cudaEvent_t copied[2] = {nullptr, nullptr};
        CK(cudaEventCreateWithFlags(&copied[0], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&copied[1], cudaEventDisableTiming));
        for (int b = 0; b < nb; ++b) {
            double* stage = hx + (size_t)(b & 1) * m;
            if (b >= 2) CK(cudaEventSynchronize(copied[b & 1]));   // the staging half is free again
            if (m) memcpy(stage, xs + (size_t)b * m, sizeof(double) * m);
            if (m) CK(cudaMemcpyAsync(h->d_x, stage, sizeof(double) * m, cudaMemcpyHostToDevice, h->stream));
            CK(cudaEventRecord(copied[b & 1], h->stream));
            if (m > 0) {
                k_edge_weights<<<h->grid_for(m), kBlock, 0, h->stream>>>(m, h->d_x, h->d_kappa, min_sel_tol, h->d_ew + h->nf);
                h->c_launches++;
            }
            launch_assemble(h, false);
            h->have_x = true;
            enqueue_fiedler_device(h, tol, max_steps, false);
            CK(cudaMemcpyAsync(hs + b, h->d_sc, sizeof(LzScalars), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaMemcpyAsync(hr + b, h->d_rr_out, sizeof(RrOut), cudaMemcpyDeviceToHost, h->stream));
        }
        CK(cudaStreamSynchronize(h->stream));
        std::vector<int> redo;
        for (int b = 0; b < nb; ++b) {
            const LzScalars& sc = hs[b];
            const RrOut& r = hr[b];
            h->c_steps += r.phases;
            h->c_spmv += r.phases;
            const double res = (sc.lnorm > 0.0 && sc.vv > 0.0) ? sc.res1 / (std::sqrt(sc.vv) * sc.lnorm) : 1.0;
            if (!(sc.lnorm > 0.0) || r.status <= 0 || r.k <= 0 || !(res < tol)) {
                redo.push_back(b);
                continue;
            }
            lambda2[b] = sc.vLv / sc.vv;
            if (resid) resid[b] = res;
        }
        cudaEventDestroy(copied[0]);
        cudaEventDestroy(copied[1]);
        cudaFreeHost(hs);
        cudaFreeHost(hr);
        cudaFreeHost(hx);
        for (int b : redo) {
            h->c_dev_fallbacks++;
            h->force_host_rr = true;
            one_by_one(b);
            h->force_host_rr = false;
        }
        h->have_g = false;
        return status;
    });
}

int macb_gradient(macb_handle h, double* g) {
    return guarded(h, [&]() {
        launch_gradient(h);
        if (g && h->m) {
            PhaseTimer pt(h, MACB_T_COPY);
            CK(cudaMemcpyAsync(g, h->d_g, sizeof(double) * h->m, cudaMemcpyDeviceToHost, h->stream));
        }
        CK(cudaStreamSynchronize(h->stream));
        return (int)MACB_OK;
    });
}

int macb_topk(macb_handle h, int64_t k, double* s) {
    return guarded(h, [&]() {
        if (!h->have_g) throw ArgFail{"macb_topk: call macb_gradient first", MACB_ERR_STATE};
        if (k < 0 || k > h->m) throw ArgFail{"macb_topk: k out of range", MACB_ERR_ARG};
        launch_topk(h, h->d_g, h->d_x, k, h->d_sel);
        h->have_sel = true;
        if (s && h->m) {
            k_mask_to_double<<<h->grid_for(h->m), kBlock, 0, h->stream>>>(h->m, h->d_sel, h->d_tmp_m);
            h->c_launches++;
            CK(cudaMemcpyAsync(s, h->d_tmp_m, sizeof(double) * h->m, cudaMemcpyDeviceToHost, h->stream));
        }
        CK(cudaStreamSynchronize(h->stream));
        return (int)MACB_OK;
    });
}

// Tie-broken nearest rounding on the device.  Stage 1: radix select on t = round(w, decimals).  If more elements
// tie with the k-th value of t than are still needed, stage 2 selects among exactly those by edge weight.
static void round_nearest_device(macb_ctx* h, int64_t k, int decimals, double* out_host) {
    const int64_t m = h->m;
    if (!h->d_tmp_m2) {
        h->d_tmp_m2 = dalloc<double>(m);
        h->d_tmp_m3 = dalloc<double>(m);
        h->d_sel_state2 = dalloc<SelState>(1);
    }
    const int grid = h->grid_for(m);
    double p10 = 1.0;
    for (int i = 0; i < decimals; ++i) p10 *= 10.0;
    // d_tmp_m holds w (uploaded by the caller); t -> d_tmp_m2
    k_round_decimals<<<grid, kBlock, 0, h->stream>>>(m, h->d_tmp_m, p10, h->d_tmp_m2);
    launch_topk(h, h->d_tmp_m2, h->d_x, k, h->d_sel, h->d_sel_state);   // leaves the k-th key of t in d_sel_state
    const bool ties = k > 0 && h->h_sel_state->eq_total != h->h_sel_state->remaining;
    if (!ties) {
        k_mask_to_double<<<grid, kBlock, 0, h->stream>>>(m, h->d_sel, h->d_tmp_m3);
        h->c_launches += 2;
    } else {
        const int64_t need = h->h_sel_state->remaining;
        k_tie_keys<<<grid, kBlock, 0, h->stream>>>(m, h->d_tmp_m2, h->d_kappa, h->d_sel_state, h->d_tmp_m3);
        launch_topk(h, h->d_tmp_m3, h->d_x, need, h->d_sel, h->d_sel_state2);
        k_round_merge<<<grid, kBlock, 0, h->stream>>>(m, h->d_tmp_m2, h->d_sel_state, h->d_sel, h->d_tmp_m3);
        h->c_launches += 3;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_host, h->d_tmp_m3, sizeof(double) * m, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->have_sel = false;
}

int macb_round_nearest(macb_handle h, const double* w, int64_t k, int decimals, double* rounded) {
    return guarded(h, [&]() {
        if (!w || !rounded) throw ArgFail{"macb_round_nearest: NULL buffer", MACB_ERR_ARG};
        if (k < 0 || k > h->m || decimals < 0 || decimals > 15) throw ArgFail{"macb_round_nearest: bad k / decimals", MACB_ERR_ARG};
        if (h->m == 0) return (int)MACB_OK;
        CK(cudaMemcpyAsync(h->d_tmp_m, w, sizeof(double) * h->m, cudaMemcpyHostToDevice, h->stream));
        round_nearest_device(h, k, decimals, rounded);
        return (int)MACB_OK;
    });
}

namespace {
// Handles of the array-only entry points (no graph: the LP oracle and the rounding need only per-candidate arrays), kept per
// (device, m) so that a user-supplied Frank-Wolfe loop calling solve_subset_box_lp every iteration does not pay a handle
// (allocations, stream, pinned buffers: ~10 ms at m = 1M) per call.  At most kDenseCache entries, least recently used first out.
constexpr size_t kDenseCache = 4;
struct DenseEntry {
    int device;
    int64_t m;
    macb_handle h;
    uint64_t stamp;
};
std::mutex g_dense_mutex;
std::vector<DenseEntry> g_dense;
uint64_t g_dense_clock = 0;

// g_dense_mutex held by the caller
macb_handle dense_handle(int device, int64_t m, int* rc) {
    for (DenseEntry& e : g_dense)
        if (e.device == device && e.m == m) {
            e.stamp = ++g_dense_clock;
            *rc = MACB_OK;
            return e.h;
        }
    std::vector<int32_t> zi((size_t)m, 0);
    std::vector<double> zk((size_t)m, 0.0);
    macb_handle h = nullptr;
    *rc = macb_create(1, 0, nullptr, nullptr, nullptr, m, zi.data(), zi.data(), zk.data(), device, &h);
    if (*rc != MACB_OK) return nullptr;
    if (g_dense.size() >= kDenseCache) {
        size_t old = 0;
        for (size_t q = 1; q < g_dense.size(); ++q)
            if (g_dense[q].stamp < g_dense[old].stamp) old = q;
        macb_destroy(g_dense[old].h);
        g_dense.erase(g_dense.begin() + (long)old);
    }
    g_dense.push_back(DenseEntry{device, m, h, ++g_dense_clock});
    return h;
}
}  // namespace

void macb_dense_cache_clear(void) {
    std::lock_guard<std::mutex> lk(g_dense_mutex);
    for (DenseEntry& e : g_dense) macb_destroy(e.h);
    g_dense.clear();
}

int macb_round_nearest_dense(int device, const double* w, const double* weights, int64_t m, int64_t k, int decimals,
                             double* rounded) {
    if (!w || !weights || !rounded || m < 1 || k < 0 || k > m) {
        g_create_error = "macb_round_nearest_dense: bad arguments";
        return MACB_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(g_dense_mutex);
    int rc = MACB_OK;
    macb_handle h = dense_handle(device, m, &rc);
    if (!h) return rc;
    rc = guarded(h, [&]() {   // this call's weights in place of the handle's
        CK(cudaMemcpyAsync(h->d_kappa, weights, sizeof(double) * m, cudaMemcpyHostToDevice, h->stream));
        return (int)MACB_OK;
    });
    if (rc == MACB_OK) rc = macb_round_nearest(h, w, k, decimals, rounded);
    if (rc != MACB_OK) g_create_error = h->err;
    return rc;
}

int macb_topk_dense(int device, const double* g, int64_t m, int64_t k, double* s) {
    if (!g || !s || m < 1 || k < 0 || k > m) {
        g_create_error = "macb_topk_dense: bad arguments";
        return MACB_ERR_ARG;
    }
    std::lock_guard<std::mutex> lk(g_dense_mutex);
    int rc = MACB_OK;
    macb_handle h = dense_handle(device, m, &rc);
    if (!h) return rc;
    rc = guarded(h, [&]() {
        CK(cudaMemcpyAsync(h->d_g, g, sizeof(double) * m, cudaMemcpyHostToDevice, h->stream));
        launch_topk(h, h->d_g, h->d_x, k, h->d_sel);
        k_mask_to_double<<<h->grid_for(m), kBlock, 0, h->stream>>>(m, h->d_sel, h->d_tmp_m);
        CK(cudaMemcpyAsync(s, h->d_tmp_m, sizeof(double) * m, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return (int)MACB_OK;
    });
    if (rc != MACB_OK) g_create_error = h->err;
    return rc;
}

namespace {
void ensure_slots(macb_ctx* c) {
    if (c->slots_ready) return;
    c->slot[0].h_sc = c->h_sc; c->slot[0].h_rr = c->h_rr; c->slot[0].h_sel = c->h_sel_state;
    c->slot[0].it0 = c->it0; c->slot[0].it1 = c->it1; c->slot[0].lz0 = c->lz0; c->slot[0].lz1 = c->lz1;
    CK(cudaMallocHost(&c->slot[1].h_sc, sizeof(LzScalars)));
    CK(cudaMallocHost(&c->slot[1].h_rr, sizeof(RrOut)));
    CK(cudaMallocHost(&c->slot[1].h_sel, sizeof(SelState)));
    for (cudaEvent_t* e : {&c->slot[1].it0, &c->slot[1].it1, &c->slot[1].lz0, &c->slot[1].lz1}) CK(cudaEventCreate(e));
    CK(cudaEventCreateWithFlags(&c->slot[0].done, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->slot[1].done, cudaEventDisableTiming));
    c->slots_ready = true;
}
void use_slot(macb_ctx* c, int s) {
    c->h_sc = c->slot[s].h_sc; c->h_rr = c->slot[s].h_rr; c->h_sel_state = c->slot[s].h_sel;
    c->it0 = c->slot[s].it0; c->it1 = c->slot[s].it1; c->lz0 = c->slot[s].lz0; c->lz1 = c->slot[s].lz1;
}

// The Frank-Wolfe loop with the host OFF the critical path.  Iteration i is enqueued completely (assemble L(x_i), eigen-solve
// with the stop decision on the device, gradient, top-k, scalars -> pinned set i & 1, event); then -- BEFORE the host waits for
// that event -- the update x_{i+1} = x_i + gamma (s_i - x_i) into the other iterate buffer and the whole of iteration i+1 are
// enqueued as well.  The host then waits for iteration i's event, applies the two stopping tests of frankwolfe.py:65-74 and
// either goes on (the device never idled) or stops (x_i is intact in its buffer; the speculative work is drained).  A device-side
// decision that did not pass the residual test, or ties straddling the budget in the LP step, drain the stream and repeat
// iteration i on the synchronous path.  Measured at the headline size: 0.30 ms per iteration of host latency at 1 GPU and
// 0.79 ms with eight processes on one host -- the whole loss of the 1 -> 8 scaling curve -- go away.
int fw_run_pipelined(macb_ctx* h, int64_t k, int max_iters, double rel_gap_tol, double grad_norm_tol, double fiedler_tol,
                     double min_sel_tol, int fiedler_max_steps, int warm, double& u, int& it_out, double* f_hist, double* u_hist) {
    ensure_slots(h);
    const int64_t m = h->m;
    if (!h->d_x_alt) h->d_x_alt = dalloc<double>(m);
    double* xbuf[2] = {h->d_x, h->d_x_alt};
    const int max_steps = fiedler_max_steps > 0 ? fiedler_max_steps : 20000;
    int status = MACB_OK;
    auto enqueue_iter = [&](int it) {
        use_slot(h, it & 1);
        h->d_x = xbuf[it & 1];
        if (h->bench_time_iters) {
            if (h->bench_flush) {
                if (!h->d_flush) CK(cudaMalloc(&h->d_flush, kFlushBytes));
                CK(cudaMemsetAsync(h->d_flush, it & 0xff, kFlushBytes, h->stream));
            }
            CK(cudaEventRecord(h->it0, h->stream));
        }
        launch_assemble(h, false);  // L(x)                   mac.py:115 -> :74
        h->have_x = true;
        enqueue_fiedler_device(h, fiedler_tol, max_steps, warm && it > 0);   // f, v     mac.py:115 -> fiedler.py:9
        launch_gradient(h);         // g                      mac.py:117-124
        launch_topk(h, h->d_g, h->d_x, k, h->d_sel, nullptr, true);  // s    frankwolfe.py:58
        CK(cudaMemcpyAsync(h->h_sc, h->d_sc, sizeof(LzScalars), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(h->h_rr, h->d_rr_out, sizeof(RrOut), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaEventRecord(h->slot[it & 1].done, h->stream));
    };
    auto enqueue_update = [&](int it) {   // x_{it+1} into the other buffer; closes the timing bracket of iteration it
        const double gamma = 2.0 / ((double)it + 2.0);  // frankwolfe.py:7-8,76
        if (m > 0) {
            k_fw_update<<<h->grid_for(m), kBlock, 0, h->stream>>>(m, gamma, h->d_sel, h->d_kappa, min_sel_tol, xbuf[it & 1],
                                                                   xbuf[(it + 1) & 1], h->d_ew + h->nf);
            CK(cudaGetLastError());
            h->c_launches++;
        }
        if (h->bench_time_iters) CK(cudaEventRecord(h->slot[it & 1].it1, h->stream));
    };
    int it = 0;
    u = std::numeric_limits<double>::infinity();
    if (max_iters > 0) enqueue_iter(0);
    for (; it < max_iters; ++it) {
        const bool more = it + 1 < max_iters;
        enqueue_update(it);
        if (more) enqueue_iter(it + 1);
        CK(cudaEventSynchronize(h->slot[it & 1].done));
        use_slot(h, it & 1);
        FiedlerResult fr;
        const bool ok = finish_fiedler_device(h, fiedler_tol, fr);
        const bool ties = k > 0 && m > 0 && h->h_sel_state->eq_total != h->h_sel_state->remaining;
        if (!ok || ties) {
            // drain the speculation and repeat this iteration on the synchronous path, from the intact x_it
            CK(cudaStreamSynchronize(h->stream));
            h->d_x = xbuf[it & 1];
            if (m > 0) {
                k_edge_weights<<<h->grid_for(m), kBlock, 0, h->stream>>>(m, h->d_x, h->d_kappa, min_sel_tol, h->d_ew + h->nf);
                h->c_launches++;
            }
            launch_assemble(h);
            h->have_x = true;
            if (!ok) {
                h->c_dev_fallbacks++;
                h->c_solves--;
                h->force_host_rr = true;
            }
            const int rc = run_fiedler(h, fiedler_tol, fiedler_max_steps, warm && it > 0, fr);
            h->force_host_rr = false;
            if (rc == MACB_NOT_CONVERGED) status = MACB_NOT_CONVERGED;
            launch_gradient(h);
            launch_topk(h, h->d_g, h->d_x, k, h->d_sel);   // synchronous: ranks ties by index where they straddle the budget
            CK(cudaMemcpyAsync(h->h_sc, h->d_sc, sizeof(LzScalars), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
        }
        const double f = fr.lambda2;
        u = std::min(u, f + h->h_sc->gs_minus_x);  //         frankwolfe.py:62
        if (f_hist) f_hist[it] = f;
        if (u_hist) u_hist[it] = u;
        if (h->bench_time_iters) {
            CK(cudaEventSynchronize(h->it1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, h->it0, h->it1));
            h->iter_ms.push_back(ms);
        }
        const bool stop = std::sqrt(h->h_sc->gnorm2) < grad_norm_tol    // frankwolfe.py:65
                          || (u - f) < rel_gap_tol * std::fabs(f);      // frankwolfe.py:71
        if (stop) {   // x_it is the answer: it sits untouched in its buffer
            CK(cudaStreamSynchronize(h->stream));   // drain the speculative iteration
            h->d_x = xbuf[it & 1];
            h->d_x_alt = xbuf[(it + 1) & 1];
            ++it;
            it_out = it;
            use_slot(h, 0);
            return status;
        }
        if (!ok || ties) {   // the speculation was drained and used a wrong selection: enqueue update and next iteration again
            enqueue_update(it);
            if (more) enqueue_iter(it + 1);
        }
    }
    h->d_x = xbuf[it & 1];
    h->d_x_alt = xbuf[(it + 1) & 1];
    it_out = it;
    use_slot(h, 0);
    return status;
}
}  // namespace

int macb_fw_run(macb_handle h, int64_t k, const double* x_init, int max_iters, double rel_gap_tol, double grad_norm_tol,
                double fiedler_tol, double min_sel_tol, int fiedler_max_steps, int warm, double* w, double* u_out,
                int* iters_done, double* f_hist, double* u_hist) {
    return guarded(h, [&]() {
        if (!x_init && h->m > 0) throw ArgFail{"macb_fw_run: x_init is NULL", MACB_ERR_ARG};
        if (k < 0 || k > h->m) throw ArgFail{"macb_fw_run: k out of range", MACB_ERR_ARG};
        {
            PhaseTimer pt(h, MACB_T_COPY);
            if (h->m) CK(cudaMemcpyAsync(h->d_x, x_init, sizeof(double) * h->m, cudaMemcpyHostToDevice, h->stream));
        }
        h->min_sel_tol = min_sel_tol;
        if (h->m > 0) {
            k_edge_weights<<<h->grid_for(h->m), kBlock, 0, h->stream>>>(h->m, h->d_x, h->d_kappa, min_sel_tol, h->d_ew + h->nf);
            h->c_launches++;
        }
        double u = std::numeric_limits<double>::infinity();
        int status = MACB_OK;
        int it = 0;
        h->iter_ms.clear();
        auto close_iter = [&]() {
            if (!h->bench_time_iters) return;
            CK(cudaEventRecord(h->it1, h->stream));
            CK(cudaEventSynchronize(h->it1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, h->it0, h->it1));
            h->iter_ms.push_back(ms);
        };
        if (h->n >= 2 && max_iters > 0) ensure_basis(h, fiedler_max_steps > 0 ? fiedler_max_steps : 20000);
        if (max_iters > 0 && h->d_basis && device_fiedler_available(h) && !h->profile && !getenv("MACB_FW_SYNC")) {
            status = fw_run_pipelined(h, k, max_iters, rel_gap_tol, grad_norm_tol, fiedler_tol, min_sel_tol, fiedler_max_steps, warm, u, it,
                                      f_hist, u_hist);
            max_iters = 0;   // skip the synchronous loop below
        }
        for (; it < max_iters; ++it) {
            if (h->bench_time_iters) {
                if (h->bench_flush) {
                    if (!h->d_flush) CK(cudaMalloc(&h->d_flush, kFlushBytes));
                    CK(cudaMemsetAsync(h->d_flush, it & 0xff, kFlushBytes, h->stream));
                }
                CK(cudaEventRecord(h->it0, h->stream));
            }
            FiedlerResult fr;
            int rc = MACB_OK;
            if (h->d_basis && device_fiedler_available(h)) {
                // The whole iteration is enqueued without a single read-back -- assemble L(x), eigen-solve with the stop
                // decision on the device, gradient, top-k -- and the host synchronises ONCE to read its scalars.
                launch_assemble(h, false);  // L(x)                   mac.py:115 -> :74
                h->have_x = true;
                {
                    PhaseTimer pt(h, MACB_T_FIEDLER);
                    enqueue_fiedler_device(h, fiedler_tol, fiedler_max_steps > 0 ? fiedler_max_steps : 20000, warm && it > 0);
                }
                launch_gradient(h);         // g                      mac.py:117-124
                launch_topk(h, h->d_g, h->d_x, k, h->d_sel, nullptr, true);  // s    frankwolfe.py:58
                CK(cudaMemcpyAsync(h->h_sc, h->d_sc, sizeof(LzScalars), cudaMemcpyDeviceToHost, h->stream));
                CK(cudaMemcpyAsync(h->h_rr, h->d_rr_out, sizeof(RrOut), cudaMemcpyDeviceToHost, h->stream));
                CK(cudaStreamSynchronize(h->stream));
                if (!finish_fiedler_device(h, fiedler_tol, fr)) {
                    // the device-side decision did not deliver a pair that passes the residual test: host-driven solve
                    h->c_dev_fallbacks++;
                    h->c_solves--;
                    h->force_host_rr = true;
                    rc = run_fiedler(h, fiedler_tol, fiedler_max_steps, warm && it > 0, fr);
                    h->force_host_rr = false;
                    launch_gradient(h);
                    launch_topk(h, h->d_g, h->d_x, k, h->d_sel, nullptr, true);
                    CK(cudaMemcpyAsync(h->h_sc, h->d_sc, sizeof(LzScalars), cudaMemcpyDeviceToHost, h->stream));
                    CK(cudaStreamSynchronize(h->stream));
                }
            } else {
                launch_assemble(h);  // L(x)                          mac.py:115 -> :74
                h->have_x = true;
                //                      f, v                          mac.py:115 -> fiedler.py:9
                rc = run_fiedler(h, fiedler_tol, fiedler_max_steps, warm && it > 0, fr);
                launch_gradient(h);  // g                             mac.py:117-124
                launch_topk(h, h->d_g, h->d_x, k, h->d_sel, nullptr, true);  // s    frankwolfe.py:58
                CK(cudaMemcpyAsync(h->h_sc, h->d_sc, sizeof(LzScalars), cudaMemcpyDeviceToHost, h->stream));
                CK(cudaStreamSynchronize(h->stream));
            }
            if (rc == MACB_NOT_CONVERGED) status = MACB_NOT_CONVERGED;
            if (topk_fixup(h, h->d_g, h->d_x, k, h->d_sel)) {   // ties at the k-th value: lowest index first
                CK(cudaMemcpyAsync(h->h_sc, h->d_sc, sizeof(LzScalars), cudaMemcpyDeviceToHost, h->stream));
                CK(cudaStreamSynchronize(h->stream));
            }
            const double f = fr.lambda2;
            u = std::min(u, f + h->h_sc->gs_minus_x);  //         frankwolfe.py:62
            if (f_hist) f_hist[it] = f;
            if (u_hist) u_hist[it] = u;
            if (std::sqrt(h->h_sc->gnorm2) < grad_norm_tol) {  // frankwolfe.py:65
                close_iter();
                ++it;
                break;
            }
            if ((u - f) < rel_gap_tol * std::fabs(f)) {  //       frankwolfe.py:71
                close_iter();
                ++it;
                break;
            }
            {
                PhaseTimer pt(h, MACB_T_UPDATE);
                const double gamma = 2.0 / ((double)it + 2.0);  // frankwolfe.py:7-8,76
                if (h->m > 0) {
                    k_fw_update<<<h->grid_for(h->m), kBlock, 0, h->stream>>>(h->m, gamma, h->d_sel, h->d_kappa, min_sel_tol,
                                                                            h->d_x, h->d_x, h->d_ew + h->nf);
                    CK(cudaGetLastError());
                    h->c_launches++;
                }
            }
            close_iter();
        }
        // the loop leaves L(x) stale with respect to d_x if it ran to max_iters; mark state accordingly
        h->have_x = false;
        h->have_v = false;
        h->have_g = false;
        if (iters_done) *iters_done = it;
        if (u_out) *u_out = u;
        if (w && h->m) {
            PhaseTimer pt(h, MACB_T_COPY);
            CK(cudaMemcpyAsync(w, h->d_x, sizeof(double) * h->m, cudaMemcpyDeviceToHost, h->stream));
        }
        CK(cudaStreamSynchronize(h->stream));
        if (status == MACB_NOT_CONVERGED) h->err = "macb_fw_run: an eigen-solve did not reach tol within fiedler_max_steps";
        return status;
    });
}

int macb_counters(macb_handle h, int64_t* kernel_launches, int64_t* spmv_launches, int64_t* lanczos_steps,
                  int64_t* fiedler_solves, double* phase_ms) {
    if (!h) return MACB_ERR_ARG;
    if (kernel_launches) *kernel_launches = h->c_launches;
    if (spmv_launches) *spmv_launches = h->c_spmv;
    if (lanczos_steps) *lanczos_steps = h->c_steps;
    if (fiedler_solves) *fiedler_solves = h->c_solves;
    if (phase_ms)
        for (int i = 0; i < MACB_T_COUNT; ++i) phase_ms[i] = h->phase_ms[i];
    return MACB_OK;
}

int macb_reset_counters(macb_handle h) {
    if (!h) return MACB_ERR_ARG;
    h->c_launches = h->c_spmv = h->c_steps = h->c_solves = 0;
    h->lz_kernel_ms = 0.0;
    h->lz_kernel_phases = 0;
    h->lz_algo_bytes = 0.0;
    h->c_dev_fallbacks = 0;
    for (int i = 0; i < MACB_T_COUNT; ++i) h->phase_ms[i] = 0.0;
    return MACB_OK;
}

int macb_set_profile(macb_handle h, int on) {
    if (!h) return MACB_ERR_ARG;
    h->profile = on != 0;
    return MACB_OK;
}

int macb_sizes(macb_handle h, int64_t* n, int64_t* m, int64_t* nnz_union, int64_t* nnz_active) {
    if (!h) return MACB_ERR_ARG;
    if (n) *n = h->n;
    if (m) *m = h->m;
    if (nnz_union) *nnz_union = h->nnz;
    if (nnz_active) *nnz_active = h->nnz_active;
    return MACB_OK;
}

int macb_l2_flush(macb_handle h) {
    return guarded(h, [&]() {
        if (!h->d_flush) CK(cudaMalloc(&h->d_flush, kFlushBytes));
        CK(cudaMemsetAsync(h->d_flush, 0xA5, kFlushBytes, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        return (int)MACB_OK;
    });
}

int macb_set_bench(macb_handle h, int time_iters, int flush_l2_between_iters) {
    if (!h) return MACB_ERR_ARG;
    h->bench_time_iters = time_iters != 0;
    h->bench_flush = flush_l2_between_iters != 0;
    return MACB_OK;
}

int macb_iter_ms(macb_handle h, double* ms, int cap, int* count) {
    if (!h) return MACB_ERR_ARG;
    int nrec = (int)h->iter_ms.size();
    if (count) *count = nrec;
    if (ms)
        for (int i = 0; i < std::min(cap, nrec); ++i) ms[i] = h->iter_ms[i];
    return MACB_OK;
}

#ifdef MACB_PTIMING
extern "C" int macb_debug_ptiming(macb_handle h, long long* out /*[64][ncta][4] + [64][ncta]*/, int* ncta) {
    return guarded(h, [&]() {
        if (ncta) *ncta = h->p_ncta;
        if (out && h->d_ptiming)
            CK(cudaMemcpy(out, h->d_ptiming, sizeof(long long) * 64 * h->p_ncta * 9, cudaMemcpyDeviceToHost));
        return (int)MACB_OK;
    });
}
#endif

int macb_lanczos_kernel_time(macb_handle h, double* ms, int64_t* phases, double* algo_bytes_per_phase) {
    if (!h) return MACB_ERR_ARG;
    if (ms) *ms = h->lz_kernel_ms;
    if (phases) *phases = h->lz_kernel_phases;
    // one Lanczos phase = one SpMV ((nnz + n) * 12 + 4 (n + 1) + 16 n) plus what the engine writes per node:
    // the 32-byte state sector and the 8-byte basis entry (sector engines), or the new z and u entries and the basis entry
    // (k_lanczos_pipe)
    if (algo_bytes_per_phase && h->lz_algo_bytes > 0.0 && h->lz_kernel_phases > 0) {
        // device-side decision: bytes follow the ACTIVE slots of every timed launch (zero-weight slots issue no gather and no
        // weight load): sum over launches of phases x [(nnz_active + n) 12 + 4 (n + 1) + 16 n + 24 n] / phases
        *algo_bytes_per_phase = h->lz_algo_bytes / (double)h->lz_kernel_phases;
    } else if (algo_bytes_per_phase) {
        const double spmv = (double)(h->nnz + h->n) * 12.0 + ((double)h->n + 1.0) * 4.0 + 16.0 * (double)h->n;
        const bool vec = h->persist_v == 5;
        *algo_bytes_per_phase = spmv + (vec ? 24.0 : 40.0) * (double)h->n;
    }
    return MACB_OK;
}

const char* macb_lanczos_kernel_name(macb_handle h) {
    if (!h || !h->d_basis) return "";
    switch (h->persist_v) {
        case 5: return "k_lanczos_pipe";
        case 4: return "k_lanczos_small2";
        case 3: return "k_lanczos_slots";
        default: return "k_lanczos_persist";
    }
}

int macb_lanczos_footprint(macb_handle h, int32_t* ctas, int32_t* sm_count) {
    if (!h || !ctas || !sm_count) return MACB_ERR_ARG;
    return guarded(h, [&]() {
        ensure_basis(h, 0);   // builds the engine (layout, buffers) if no solve has done so yet
        *ctas = h->p_ncta > 0 ? h->p_ncta + (h->dev_rr ? 1 : 0) : 0;
        *sm_count = h->sm_count;
        return (int)MACB_OK;
    });
}

int macb_measure_l2_bandwidth(int device, int64_t bytes, int reps, double* gbs) {
    if (bytes < (1 << 20) || reps < 1 || !gbs) return MACB_ERR_ARG;
    try {
        if (device >= 0) CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        int dev = 0;
        CK(cudaGetDevice(&dev));
        CK(cudaGetDeviceProperties(&prop, dev));
        const int64_t n16 = bytes / 16;
        double2* buf = nullptr;
        double* sink = nullptr;
        CK(cudaMalloc(&buf, (size_t)n16 * 16));
        CK(cudaMalloc(&sink, 8));
        CK(cudaMemset(buf, 0, (size_t)n16 * 16));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        const int grid = prop.multiProcessorCount;
        k_l2_read<<<grid, 1024>>>(buf, n16, 2, sink);   // warm-up: brings the buffer into L2
        CK(cudaEventRecord(e0));
        k_l2_read<<<grid, 1024>>>(buf, n16, reps, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        *gbs = (double)grid * reps * (double)(n16 * 16) / ((double)ms * 1e-3) / 1e9;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaFree(buf);
        cudaFree(sink);
        return MACB_OK;
    } catch (const CudaFail& e) {
        char b[256];
        snprintf(b, sizeof(b), "macb_measure_l2_bandwidth: CUDA error %d (%s)", (int)e.e, cudaGetErrorString(e.e));
        g_create_error = b;
        cudaGetLastError();
        return MACB_ERR_CUDA;
    }
}

int macb_device_rr_stats(macb_handle h, int* enabled, int64_t* fallbacks, int* last_status, int* last_k, int* last_checks,
                         double* last_theta_est_target /*[3], may be NULL*/) {
    if (!h) return MACB_ERR_ARG;
    if (enabled) *enabled = device_fiedler_available(h) ? 1 : 0;
    if (fallbacks) *fallbacks = h->c_dev_fallbacks;
    if (last_status) *last_status = h->h_rr ? h->h_rr->status : 0;
    if (last_k) *last_k = h->h_rr ? h->h_rr->k : 0;
    if (last_checks) *last_checks = h->h_rr ? h->h_rr->checks : 0;
    if (last_theta_est_target && h->h_rr) {
        last_theta_est_target[0] = h->h_rr->theta;
        last_theta_est_target[1] = h->h_rr->est;
        last_theta_est_target[2] = h->h_rr->target;
        last_theta_est_target[3] = (double)h->h_rr->cyc_wait;
        last_theta_est_target[4] = (double)h->h_rr->cyc_compute;
        last_theta_est_target[5] = (double)h->h_rr->lag;
        last_theta_est_target[6] = (double)h->h_rr->rounds;
        for (int i = 0; i < 6; ++i) last_theta_est_target[7 + i] = (double)h->h_rr->cyc_stage[i];
    }
    return MACB_OK;
}

int macb_device_sync(macb_handle h) {
    return guarded(h, [&]() {
        CK(cudaDeviceSynchronize());
        return (int)MACB_OK;
    });
}

// Builds the chunked jagged-diagonal copy of the pattern for k_spmv_jds (kernels.cuh).  Returns false when a row is
// longer than a chunk can hold (the CSR kernel then stays in charge).
static bool build_spmv_jds(macb_ctx* c) {
    const int n = c->n;
    const int64_t nnz = c->nnz;
    std::vector<int> col((size_t)nnz), eid((size_t)nnz);
    if (nnz) {
        CK(cudaMemcpy(col.data(), c->d_col, sizeof(int) * nnz, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(eid.data(), c->d_eid, sizeof(int) * nnz, cudaMemcpyDeviceToHost));
    }
    const std::vector<int32_t>& rp = c->h_rp;
    // pass A: chunk boundaries; can every chunk use 16-bit column offsets?
    std::vector<int> chunk_row{0};
    std::vector<int64_t> chunk_slot{0};
    bool col16 = !getenv("MACB_SPMV_COL32");
    for (int r = 0; r < n;) {
        int e = r;
        while (e < n && e - r < kSjRows && (int64_t)rp[e + 1] - rp[r] <= kSjCap) ++e;
        if (e == r) return false;   // a single row exceeds the chunk capacity
        int cmin = std::numeric_limits<int>::max(), cmax = 0;
        for (int64_t s = rp[r]; s < rp[e]; ++s) {
            cmin = std::min(cmin, col[(size_t)s]);
            cmax = std::max(cmax, col[(size_t)s]);
        }
        if (rp[e] > rp[r] && cmax - cmin > 65535) col16 = false;
        chunk_row.push_back(e);
        chunk_slot.push_back(rp[e]);
        r = e;
    }
    const int nchunks = (int)chunk_row.size() - 1;
    std::vector<int> chunk_jd{0}, chunk_col0((size_t)nchunks, 0), jd, perm((size_t)n), len((size_t)n), jeid((size_t)nnz, 0);
    std::vector<int> jcol32(col16 ? 0 : (size_t)nnz, 0);
    std::vector<unsigned int> word((size_t)nnz, 0u);
    std::vector<int> order, cnt, base, pos2eid;
    std::vector<std::pair<int, int>> key;
    for (int q = 0; q < nchunks; ++q) {
        const int r = chunk_row[q], e = chunk_row[q + 1], R = e - r;
        const int ns = (int)(rp[e] - rp[r]);
        const int64_t s0 = chunk_slot[q];
        order.resize(R);
        for (int t = 0; t < R; ++t) order[t] = r + t;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return rp[x + 1] - rp[x] > rp[y + 1] - rp[y]; });
        const int maxlen = R ? rp[order[0] + 1] - rp[order[0]] : 0;
        if (maxlen > kSjMaxLen) return false;
        cnt.assign((size_t)maxlen + 1, 0);
        for (int t = 0; t < R; ++t) {
            perm[r + t] = order[t];
            len[r + t] = rp[order[t] + 1] - rp[order[t]];
            for (int d = 0; d < len[r + t]; ++d) cnt[d]++;
        }
        base.assign((size_t)maxlen + 1, 0);
        for (int d = 1; d <= maxlen; ++d) base[d] = base[d - 1] + cnt[d - 1];
        for (int d = 0; d <= maxlen; ++d) jd.push_back(base[d]);
        chunk_jd.push_back((int)jd.size());
        key.clear();
        pos2eid.assign((size_t)std::max(ns, 1), 0);
        int cmin = std::numeric_limits<int>::max();
        for (int t = 0; t < R; ++t) {
            const int row = order[t];
            for (int d = 0; d < len[r + t]; ++d) {
                const int cc = col[(size_t)rp[row] + d];
                key.push_back({cc, base[d] + t});
                pos2eid[(size_t)base[d] + t] = eid[(size_t)rp[row] + d];
                cmin = std::min(cmin, cc);
            }
        }
        if (ns == 0) cmin = 0;
        chunk_col0[q] = cmin;
        std::sort(key.begin(), key.end());
        for (int p = 0; p < ns; ++p) {
            if (col16) {
                word[(size_t)s0 + p] = (unsigned int)(key[p].first - cmin) | ((unsigned int)key[p].second << 16);
            } else {
                jcol32[(size_t)s0 + p] = key[p].first;
                word[(size_t)s0 + p] = (unsigned int)key[p].second;
            }
            jeid[(size_t)s0 + p] = pos2eid[(size_t)key[p].second];
        }
    }
    c->sj_nchunks = nchunks;
    c->sj_col16 = col16;
    auto up = [&](auto*& dptr, const auto& v) {
        using T = typename std::remove_const<typename std::remove_reference<decltype(v[0])>::type>::type;
        dptr = dalloc<T>(v.size());
        if (!v.empty()) CK(cudaMemcpy(dptr, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
    };
    up(c->d_sj_chunk_row, chunk_row);
    up(c->d_sj_chunk_slot, chunk_slot);
    up(c->d_sj_chunk_jd, chunk_jd);
    up(c->d_sj_col0, chunk_col0);
    up(c->d_sj_jd, jd);
    up(c->d_sj_perm, perm);
    up(c->d_sj_len, len);
    up(c->d_sj_eid, jeid);
    up(c->d_sj_word, word);
    if (!col16) up(c->d_sj_col, jcol32);
    c->d_sj_val = dalloc<double>((size_t)nnz);
    if (c->have_x) {
        k_assemble_jds<<<c->grid_for(nnz), kBlock, 0, c->stream>>>(nnz, c->d_sj_eid, c->d_ew, c->d_sj_val);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(c->stream));
    }
    return true;
}

int macb_spmv_engine(macb_handle h, int engine) {
    return guarded(h, [&]() {
        if (engine != 0 && engine != 1) throw ArgFail{"macb_spmv_engine: engine must be 0 (CSR) or 1 (chunked jagged-diagonal)", MACB_ERR_ARG};
        if (engine == 1 && !h->d_sj_val && !build_spmv_jds(h))
            throw ArgFail{"macb_spmv_engine: a row is longer than a chunk of the jagged-diagonal kernel", MACB_ERR_ARG};
        h->spmv_engine = engine;
        return (int)MACB_OK;
    });
}

int macb_spmv_bench(macb_handle h, int reps, int flush_l2, double* avg_ms, double* algo_bytes) {
    return guarded(h, [&]() {
        if (!h->have_x) throw ArgFail{"macb_spmv_bench: call macb_set_x first", MACB_ERR_STATE};
        if (reps < 1) reps = 1;
        CK(cudaMemcpyAsync(h->d_tmp_n, h->d_x0, sizeof(double) * h->n, cudaMemcpyDeviceToDevice, h->stream));
        for (int i = 0; i < 3; ++i) launch_spmv<0>(h, h->d_tmp_n, h->d_y);
        double total = 0.0;
        if (flush_l2) {
            if (!h->d_flush) CK(cudaMalloc(&h->d_flush, kFlushBytes));
            for (int i = 0; i < reps; ++i) {
                CK(cudaMemsetAsync(h->d_flush, i, kFlushBytes, h->stream));
                CK(cudaEventRecord(h->ev0, h->stream));
                launch_spmv<0>(h, h->d_tmp_n, h->d_y);
                CK(cudaEventRecord(h->ev1, h->stream));
                CK(cudaEventSynchronize(h->ev1));
                float ms = 0.f;
                CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
                total += ms;
            }
        } else {
            CK(cudaEventRecord(h->ev0, h->stream));
            for (int i = 0; i < reps; ++i) launch_spmv<0>(h, h->d_tmp_n, h->d_y);
            CK(cudaEventRecord(h->ev1, h->stream));
            CK(cudaEventSynchronize(h->ev1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
            total = ms;
        }
        h->c_launches += reps + 3;
        h->c_spmv += reps + 3;
        if (avg_ms) *avg_ms = total / reps;
        // nnz (8 + 4) + (n + 1) 4 + 8 n (x) + 8 n (y); the diagonal counts as n more non-zeros
        if (algo_bytes) *algo_bytes = (double)(h->nnz + h->n) * 12.0 + ((double)h->n + 1.0) * 4.0 + 16.0 * (double)h->n;
        return (int)MACB_OK;
    });
}

}  // extern "C"

// ================================================================================================ K-sweep farm (NCCL)
// One process per GPU; every rank holds the same graph and takes its share of the budgets; the ONLY
// communication is one ncclAllGather of fixed-size result records at the end.  NCCL is loaded at run time (dlopen), so the
// library itself links against nothing but the CUDA runtime and still loads on a box without NCCL.
struct NcclId { char bytes[128]; };   // ncclUniqueId: passed BY VALUE to ncclCommInitRank
namespace {
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
std::mutex g_nccl_mutex;

bool load_nccl(std::string& err) {
    std::lock_guard<std::mutex> lk(g_nccl_mutex);
    if (g_nccl.lib) return true;
    const char* names[] = {getenv("MACB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* nm : names) {
        if (!nm) continue;
        lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (lib) break;
    }
    if (!lib) {
        err = "NCCL not found (dlopen libnccl.so.2 failed; set MACB_NCCL_LIB)";
        return false;
    }
    NcclApi a;
    a.lib = lib;
    a.GetUniqueId = (int (*)(void*))dlsym(lib, "ncclGetUniqueId");
    a.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(lib, "ncclCommInitRank");
    a.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
    a.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    a.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy) {
        err = "NCCL library lacks an expected symbol";
        dlclose(lib);
        return false;
    }
    g_nccl = a;
    return true;
}

// longest-processing-time-first assignment of the budgets to ranks (the same rule as mac_b200/farm.py:assign)
std::vector<int> sweep_owner(const int64_t* ks, int nk, int64_t m, int nranks) {
    std::vector<double> cost(nk);
    for (int i = 0; i < nk; ++i) cost[i] = 1.0 + (double)(m - ks[i]) / (double)std::max<int64_t>(m, 1);
    std::vector<int> order(nk), owner(nk, 0);
    for (int i = 0; i < nk; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return cost[x] > cost[y]; });
    std::vector<double> load(nranks, 0.0);
    for (int i : order) {
        int r = 0;
        for (int q = 1; q < nranks; ++q)
            if (load[q] < load[r]) r = q;
        owner[i] = r;
        load[r] += cost[i];
    }
    return owner;
}
}  // namespace

struct macb_comm {
    void* comm = nullptr;
    int nranks = 1, rank = 0, device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
};

extern "C" {

int macb_comm_unique_id(char* id /*[128]*/) {
    std::string err;
    if (!id) return MACB_ERR_ARG;
    if (!load_nccl(err)) {
        g_create_error = err;
        return MACB_ERR_STATE;
    }
    NcclId u;
    memset(&u, 0, sizeof(u));
    const int rc = g_nccl.GetUniqueId(&u);
    if (rc != 0) {
        g_create_error = std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
        return MACB_ERR_CUDA;
    }
    memcpy(id, u.bytes, 128);
    return MACB_OK;
}

int macb_comm_init(int nranks, int rank, const char* id, int device, macb_comm_t* out) {
    if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) return MACB_ERR_ARG;
    *out = nullptr;
    std::string err;
    if (!load_nccl(err)) {
        g_create_error = err;
        return MACB_ERR_STATE;
    }
    macb_comm* c = new macb_comm();
    c->nranks = nranks;
    c->rank = rank;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) device = 0;
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_create_error = "macb_comm_init: cannot select the device / create a stream";
        cudaGetLastError();
        delete c;
        return MACB_ERR_CUDA;
    }
    NcclId u;
    memcpy(u.bytes, id, 128);
    const int rc = g_nccl.CommInitRank(&c->comm, nranks, u, rank);
    if (rc != 0) {
        g_create_error = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error");
        cudaStreamDestroy(c->stream);
        delete c;
        return MACB_ERR_CUDA;
    }
    *out = c;
    return MACB_OK;
}

int macb_comm_allgather(macb_comm_t c, const void* send, void* recv, int64_t bytes_per_rank) {
    if (!c || !send || !recv || bytes_per_rank < 0) return MACB_ERR_ARG;
    if (bytes_per_rank == 0) return MACB_OK;
    void *dsend = nullptr, *drecv = nullptr;
    cudaSetDevice(c->device);
    cudaError_t e = cudaMalloc(&dsend, (size_t)bytes_per_rank);
    if (e == cudaSuccess) e = cudaMalloc(&drecv, (size_t)bytes_per_rank * c->nranks);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dsend, send, (size_t)bytes_per_rank, cudaMemcpyHostToDevice, c->stream);
    int nrc = 0;
    if (e == cudaSuccess) nrc = g_nccl.AllGather(dsend, drecv, (size_t)bytes_per_rank, /* ncclUint8 */ 1, c->comm, c->stream);
    if (e == cudaSuccess && nrc == 0) e = cudaMemcpyAsync(recv, drecv, (size_t)bytes_per_rank * c->nranks, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && nrc == 0) e = cudaStreamSynchronize(c->stream);
    if (dsend) cudaFree(dsend);
    if (drecv) cudaFree(drecv);
    if (nrc != 0) {
        c->err = std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "error");
        return MACB_ERR_CUDA;
    }
    if (e != cudaSuccess) {
        c->err = std::string("macb_comm_allgather: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return MACB_ERR_CUDA;
    }
    return MACB_OK;
}

int macb_comm_destroy(macb_comm_t c) {
    if (!c) return MACB_OK;
    if (c->comm) g_nccl.CommDestroy(c->comm);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return MACB_OK;
}

const char* macb_comm_last_error(macb_comm_t c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int macb_sweep_owner(const int64_t* ks, int nk, int64_t m, int nranks, int32_t* owner) {
    if (!ks || !owner || nk < 0 || nranks < 1) return MACB_ERR_ARG;
    std::vector<int> o = sweep_owner(ks, nk, m, nranks);
    for (int i = 0; i < nk; ++i) owner[i] = o[i];
    return MACB_OK;
}

int macb_sweep(macb_handle h, macb_comm_t comm, const int64_t* ks, int nk, const double* x_inits, int max_iters,
               double rel_gap_tol, double grad_norm_tol, double fiedler_tol, double min_sel_tol, int fiedler_max_steps,
               uint8_t* rounded, double* w, double* u, double* lambda_unrounded, int32_t* iters) {
    if (!h) return MACB_ERR_ARG;
    if (!ks || nk < 0 || (!x_inits && nk > 0 && h->m > 0) || !rounded || !u || !lambda_unrounded || !iters) {
        h->err = "macb_sweep: NULL argument";
        return MACB_ERR_ARG;
    }
    const int nranks = comm ? comm->nranks : 1, rank = comm ? comm->rank : 0;
    const int64_t m = h->m;
    const std::vector<int> owner = sweep_owner(ks, nk, m, nranks);
    // fixed-size record per budget: rounded mask (m bytes, padded to 8), then w (m doubles, optional), then u, lambda2, iters
    const int64_t mpad = (m + 7) / 8 * 8;
    const int64_t rec = mpad + (w ? 8 * m : 0) + 24;
    int per_rank = 0;
    {
        std::vector<int> cnt(nranks, 0);
        for (int i = 0; i < nk; ++i) per_rank = std::max(per_rank, ++cnt[owner[i]]);
    }
    std::vector<uint8_t> mine((size_t)std::max<int64_t>(rec * per_rank, 1), 0);
    std::vector<double> wbuf((size_t)std::max<int64_t>(m, 1)), rbuf((size_t)std::max<int64_t>(m, 1));
    int slot = 0, status = MACB_OK;
    for (int i = 0; i < nk; ++i) {
        if (owner[i] != rank) continue;
        if (ks[i] < 0 || ks[i] > m) {
            h->err = "macb_sweep: budget out of range";
            return MACB_ERR_ARG;
        }
        double ui = 0.0, lam = 0.0;
        int it = 0;
        // the g2o protocol (g2o_experiment.py:306-321): MAC.solve(K, x_init, max_iters, rounding='nearest')
        if (ks[i] >= m) {   // mac.py:173-180
            std::fill(wbuf.begin(), wbuf.end(), 1.0);
            std::fill(rbuf.begin(), rbuf.end(), 1.0);
            int rc = macb_set_x(h, wbuf.data(), min_sel_tol);
            if (rc == MACB_OK) rc = macb_fiedler(h, fiedler_tol, fiedler_max_steps, 0, &lam, nullptr, nullptr, nullptr);
            if (rc < 0) return rc;
            ui = lam;
        } else {
            int rc = macb_fw_run(h, ks[i], x_inits + (size_t)i * m, max_iters, rel_gap_tol, grad_norm_tol, fiedler_tol, min_sel_tol,
                                 fiedler_max_steps, 0, wbuf.data(), &ui, &it, nullptr, nullptr);
            if (rc < 0) return rc;
            if (rc == MACB_NOT_CONVERGED) status = rc;
            rc = macb_round_nearest(h, wbuf.data(), ks[i], 10, rbuf.data());
            if (rc < 0) return rc;
            rc = macb_set_x(h, wbuf.data(), min_sel_tol);
            if (rc == MACB_OK) rc = macb_fiedler(h, fiedler_tol, fiedler_max_steps, 0, &lam, nullptr, nullptr, nullptr);
            if (rc < 0) return rc;
        }
        uint8_t* p = mine.data() + (size_t)slot * rec;
        for (int64_t e = 0; e < m; ++e) p[e] = rbuf[e] != 0.0 ? 1 : 0;
        if (w) memcpy(p + mpad, wbuf.data(), (size_t)8 * m);
        double tail[3] = {ui, lam, (double)it};
        memcpy(p + mpad + (w ? 8 * m : 0), tail, 24);
        ++slot;
    }
    std::vector<uint8_t> all;
    const uint8_t* src = mine.data();
    if (comm && nranks > 1) {
        all.resize((size_t)rec * per_rank * nranks);
        const int rc = macb_comm_allgather(comm, mine.data(), all.data(), rec * per_rank);
        if (rc != MACB_OK) {
            h->err = comm->err;
            return rc;
        }
        src = all.data();
    }
    std::vector<int> next(nranks, 0);
    for (int i = 0; i < nk; ++i) {
        const int r = owner[i];
        const uint8_t* p = src + ((size_t)(nranks > 1 && comm ? r : 0) * per_rank + next[r]++) * rec;
        memcpy(rounded + (size_t)i * m, p, (size_t)m);
        if (w) memcpy(w + (size_t)i * m, p + mpad, (size_t)8 * m);
        double tail[3];
        memcpy(tail, p + mpad + (w ? 8 * m : 0), 24);
        u[i] = tail[0];
        lambda_unrounded[i] = tail[1];
        iters[i] = (int32_t)tail[2];
    }
    return status;
}

}  // extern "C"
