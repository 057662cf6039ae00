// sm_100a device kernels for the MAC Frank-Wolfe hot path.  All arithmetic is IEEE double.
//
// Data layout in HBM (all arrays live for the lifetime of a handle):
//   pattern   row_ptr[n+1], col[nnz], eid[nnz]    union pattern of L(w), off-diagonals only
//   values    val[nnz] (>= 0, the edge weight w_e), diag[n] (weighted degree)   rewritten per FW step
//   edges     ew[nf+m]  current weight of every edge (fixed: constant; candidate k: x_k kappa_k)
//             ci[m], cj[m], kappa[m], x[m], g[m], sel[m] (u8)
//   Lanczos   basis[(cap+1) * ld]  un-normalised Lanczos vectors u_j (row j), ld = n padded to 32
//             alpha[cap], beta[cap+1] (beta[j] = ||u_j||), ysum[cap]
// L(w) v is evaluated as  y_i = diag_i v_i - sum_s val_s v[col_s]  (the Laplacian of the reference,
// graphs.py:77-96, with the duplicate-summed diagonal kept as a separate array).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace macb {

constexpr int kBlock = 256;
constexpr int kWarpsPerBlock = kBlock / 32;

// Scalars shared between the kernels of one eigen-solve (device memory).
struct LzScalars {
    int step;          // index j of the Lanczos vector the next A-kernel multiplies
    int pad;
    double lnorm;      // max_i sum_j |L_ij|  (nx:229)
    int64_t nnz_active;
    // finalisation
    double ritz_sum, ritz_sq;   // sum y, sum y^2 of the raw Ritz vector
    double vLv, vv;             // Rayleigh quotient pieces
    double res1;                // ||L v - theta v||_1
    // gradient / LP
    double gnorm2, gdotx, gs_minus_x;
    int64_t nsel;
    double shift;               // trace(L)/n: the spectral shift of the pipelined Lanczos kernel (k_assemble)
};

struct ReduceWS {
    double* partials;        // [grid_max * 4]
    unsigned int* counter;   // zero between kernels
};

__device__ __forceinline__ double ld_nc(const double* p) { return __ldg(p); }
__device__ __forceinline__ int ld_nc(const int* p) { return __ldg(p); }

// Streaming (read-once-per-kernel) loads: keep them out of L1 so the gathered vector stays there.
__device__ __forceinline__ double ld_stream(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int ld_stream(const int* p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ double safe_inv(double b) { return (b > 1e-290) ? 1.0 / b : 0.0; }

// Coefficients of the next Lanczos vector from the four global sums (p1 = u.z, p2 = sum z, p3 = u.u, p4 = sum u).
// The chain is on the critical path of every step on every CTA, so it avoids double-precision divide and sqrt
// (each a ~100-instruction sequence): one rsqrt, then multiplications only.
struct LzCoef {
    double alpha, beta, binv, k1, k2, k3, k4;
};
__device__ __forceinline__ LzCoef lz_coefficients(double P1, double P2, double P3, double P4, double binv_prev,
                                                  double usum_prev, double inv_n) {
    LzCoef c;
    c.binv = (P3 > 1e-290) ? rsqrt(P3) : 0.0;
    c.beta = P3 * c.binv;
    c.alpha = P1 * c.binv * c.binv;
    c.k1 = c.binv;
    c.k2 = -c.alpha * c.binv;
    c.k3 = -c.beta * binv_prev;            // binv_prev = 0 on the first step
    c.k4 = -(c.k1 * P2 + c.k2 * P4 + c.k3 * usum_prev) * inv_n;
    return c;
}

// ---- block / grid reductions (fixed tree => bitwise reproducible for a fixed grid) ---------------
template <int NV, bool MAXOP = false>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* sm /*[NV * kWarpsPerBlock]*/) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double t = __shfl_xor_sync(0xffffffffu, v[i], o);
            v[i] = MAXOP ? fmax(v[i], t) : v[i] + t;
        }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < NV; ++i) sm[i * kWarpsPerBlock + warp] = v[i];
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double x = (lane < kWarpsPerBlock) ? sm[i * kWarpsPerBlock + lane] : (MAXOP ? -1.0e300 : 0.0);
#pragma unroll
            for (int o = kWarpsPerBlock / 2; o > 0; o >>= 1) {
                double t = __shfl_xor_sync(0xffffffffu, x, o);
                x = MAXOP ? fmax(x, t) : x + t;
            }
            v[i] = x;
        }
    }
}

// Every block publishes its partial; the last block to arrive reduces all partials in a fixed
// order.  Returns true on thread 0 of that last block, with the grand totals in v.
template <int NV, bool MAXOP = false>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], ReduceWS ws, double* sm, int* sm_flag) {
    block_reduce<NV, MAXOP>(v, sm);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) ws.partials[(size_t)blockIdx.x * NV + i] = v[i];
        __threadfence();
        unsigned int t = atomicAdd(ws.counter, 1u);
        *sm_flag = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!*sm_flag) return false;
    __threadfence();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = MAXOP ? -1.0e300 : 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double t = __ldcg(ws.partials + (size_t)b * NV + i);
            v[i] = MAXOP ? fmax(v[i], t) : v[i] + t;
        }
    block_reduce<NV, MAXOP>(v, sm);
    if (threadIdx.x == 0) {
        *ws.counter = 0u;
        return true;
    }
    return false;
}

// ---- K2: assembly of L(x) on the fixed pattern ----------------------------------------------------
// ew[nf + k] = x_k kappa_k if x_k > tol else 0   (mac.py:85-86)
__global__ void __launch_bounds__(kBlock) k_edge_weights(int64_t m, const double* __restrict__ x,
                                                         const double* __restrict__ kappa, double tol,
                                                         double* __restrict__ ew_cand) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        double xi = x[i];
        ew_cand[i] = (xi > tol) ? xi * kappa[i] : 0.0;
    }
}

// W lanes per row: val[s] = ew[eid[s]], diag[row] = sum_s val[s]; Lnorm = 2 max diag; active-slot count.
template <int W>
__global__ void __launch_bounds__(kBlock) k_assemble(int n, const int* __restrict__ rp, const int* __restrict__ eid,
                                                     const double* __restrict__ ew, double* __restrict__ val,
                                                     double* __restrict__ diag, LzScalars* sc, ReduceWS ws) {
    __shared__ double sm[2 * kWarpsPerBlock];
    __shared__ int flag;
    const int lane = threadIdx.x & (W - 1);
    const int sub = (blockIdx.x * kBlock + threadIdx.x) / W;
    const int nsub = gridDim.x * kBlock / W;
    double dmax = 0.0, cnt = 0.0, dsum = 0.0;
    const int rows_per_warp = 32 / W;
    for (int base = sub - (sub % rows_per_warp); base < n; base += nsub) {
        int row = base + (sub % rows_per_warp);
        double acc = 0.0;
        if (row < n) {
            int s0 = rp[row], s1 = rp[row + 1];
            for (int s = s0 + lane; s < s1; s += W) {
                double w = ew[ld_stream(eid + s)];
                val[s] = w;
                acc += w;
                cnt += (w != 0.0) ? 1.0 : 0.0;
            }
        }
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (row < n && lane == 0) {
            diag[row] = acc;
            dmax = fmax(dmax, acc);
            dsum += acc;
        }
    }
    double v[1] = {dmax};
    bool last_max = grid_reduce<1, true>(v, ws, sm, &flag);
    if (last_max) sc->lnorm = 2.0 * v[0];
    // second reduction (count) reuses the workspace after the first has fully completed in this block
    __syncthreads();
    double c[2] = {cnt, dsum};
    ReduceWS ws2 = {ws.partials + (size_t)gridDim.x, ws.counter + 1};
    bool last_cnt = grid_reduce<2, false>(c, ws2, sm, &flag);
    if (last_cnt) {
        sc->nnz_active = (int64_t)(c[0] + 0.5);
        sc->shift = c[1] / (double)n;
    }
}

// ---- K1: CSR SpMV, W lanes per row, fused reductions ---------------------------------------------
// MODE 0: y = L x
// MODE 2: Rayleigh: y = L x; vLv = x.y; vv = x.x
struct SpmvArgs {
    int n;
    int ld;
    const int* rp;
    const int* col;
    const double* val;
    const double* diag;
    const double* x;
    double* y;
    LzScalars* sc;
    ReduceWS ws;
};

template <int W, int MODE>
__global__ void __launch_bounds__(kBlock) k_spmv(SpmvArgs a) {
    __shared__ double sm[2 * kWarpsPerBlock];
    __shared__ int flag;
    const int lane = threadIdx.x & (W - 1);
    const int sub = (blockIdx.x * kBlock + threadIdx.x) / W;
    const int nsub = gridDim.x * kBlock / W;
    constexpr int rows_per_warp = 32 / W;
    const int sub_in_warp = sub % rows_per_warp;

    const double* __restrict__ x = a.x;
    const int* __restrict__ col = a.col;
    const double* __restrict__ val = a.val;

    double r0 = 0.0, r1 = 0.0;
    for (int base = sub - sub_in_warp; base < a.n; base += nsub) {
        const int row = base + sub_in_warp;
        double acc0 = 0.0, acc1 = 0.0;
        if (row < a.n) {
            const int s0 = a.rp[row], s1 = a.rp[row + 1];
            int s = s0 + lane;
            for (; s + W < s1; s += 2 * W) {
                int c0 = ld_stream(col + s), c1 = ld_stream(col + s + W);
                double v0 = ld_stream(val + s), v1 = ld_stream(val + s + W);
                acc0 = fma(v0, ld_nc(x + c0), acc0);
                acc1 = fma(v1, ld_nc(x + c1), acc1);
            }
            if (s < s1) acc0 = fma(ld_stream(val + s), ld_nc(x + ld_stream(col + s)), acc0);
        }
        acc0 += acc1;
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
        if (row < a.n && lane == 0) {
            const double xi = ld_nc(x + row);
            const double yi = a.diag[row] * xi - acc0;
            a.y[row] = yi;
            if (MODE == 2) {
                r0 = fma(xi, yi, r0);
                r1 = fma(xi, xi, r1);
            }
        }
    }
    if (MODE != 0) {
        double v[2] = {r0, r1};
        if (grid_reduce<2>(v, a.ws, sm, &flag)) {
            a.sc->vLv = v[0];
            a.sc->vv = v[1];
        }
    }
}

// ---- K1, chunked jagged-diagonal form (HBM-bound matrices) -----------------------------------------------------
// k_spmv above spends one L1TEX wavefront per non-zero on the gather of x (32 lanes, 32 different 128-byte lines) and is
// therefore bound by the SM's memory front end (~0.9 nnz/clk/SM = 3.1 TB/s of matrix stream), not by HBM.  Here the
// matrix is cut into chunks of <= kSjRows rows / <= kSjCap slots; inside a chunk the slots are stored in COLUMN order,
// so that on matrices with locality (pose graphs: |i - j| bounded) the 32 lanes of a gather share one or two lines, and
// every slot carries the position of its product in the chunk's jagged-diagonal buffer (rows sorted by decreasing length,
// diagonal d = the d-th product of every row that has one):
//   pass 1   prod[pos_s] = val_s * x[col_s]     coalesced streams: one 32-bit word (16-bit column offset | 16-bit position)
//            + the weight = 12 bytes per slot when the chunk's columns span < 65536, 14 bytes (32-bit column) otherwise;
//            eight slots in flight per thread
//   pass 2   y[row_t] = diag * x[row_t] - sum_d prod[jd[d] + t]      conflict-free shared-memory reads
// Four CTAs per SM so that one CTA's streams overlap the other's row sums.  (Staging whole chunks through the TMA unit
// -- cp.async.bulk + mbarrier, one chunk ahead -- was measured slower: the stream was never the problem, the serial
// pass 1 -> pass 2 structure of a single resident CTA is.)
constexpr int kSjBlock = 256;
constexpr int kSjMinBlocks = 5;  // 48 registers, 40 warps/SM: 62 % of measured HBM peak (4 -> 54 %, 6/8 -> 62/61 %)
constexpr int kSjRows = 256;
constexpr int kSjCap = 2048;     // slots per chunk: 16 KB of products
constexpr int kSjMaxLen = 1023;  // longest row a chunk can hold (its diagonal starts live in shared memory)
constexpr int kSjBatch = 8;

struct SpmvJdsArgs {
    int nchunks;
    const int* chunk_row;       // [nchunks + 1] first (chunk-ordered) row of every chunk
    const int64_t* chunk_slot;  // [nchunks + 1] first slot
    const int* chunk_jd;        // [nchunks + 1] first entry of the chunk's diagonal starts in jd
    const int* chunk_col0;      // [nchunks] smallest column of the chunk (COL16: col = col0 + low 16 bits)
    const int* jd;
    const int* perm;            // [n] chunk-ordered row -> row of the caller
    const int* len;             // [n] its number of slots
    const unsigned int* word;   // [slots] COL16: column offset | position << 16;  else: position
    const int* col;             // [slots] 32-bit column (!COL16 only)
    const double* val;          // [slots]
    const double* diag;         // [n] caller numbering
    const double* x;
    double* y;
};

__device__ __forceinline__ unsigned int ld_stream(const unsigned int* p) {
    unsigned int r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

template <bool COL16>
__global__ void __launch_bounds__(kSjBlock, kSjMinBlocks) k_spmv_jds(SpmvJdsArgs a) {
    __shared__ double prod[kSjCap];
    __shared__ int sjd[kSjMaxLen + 1];
    const int tid = (int)threadIdx.x;
    for (int q = blockIdx.x; q < a.nchunks; q += gridDim.x) {
        const int r0 = a.chunk_row[q], nrows = a.chunk_row[q + 1] - r0;
        const int64_t s0 = a.chunk_slot[q];
        const int ns = (int)(a.chunk_slot[q + 1] - s0);
        const int j0 = a.chunk_jd[q], nj = a.chunk_jd[q + 1] - j0;
        const int col0 = COL16 ? a.chunk_col0[q] : 0;
        for (int i = tid; i < nj; i += kSjBlock) sjd[i] = a.jd[j0 + i];
        int row = 0, len = 0;
        double dxr = 0.0;
        if (tid < nrows) {
            row = a.perm[r0 + tid];
            len = a.len[r0 + tid];
            dxr = a.diag[row] * ld_nc(a.x + row);
        }
        const unsigned int* __restrict__ word = a.word + s0;
        const int* __restrict__ col = a.col + s0;
        const double* __restrict__ val = a.val + s0;
        for (int b0 = tid; b0 < ns; b0 += kSjBatch * kSjBlock) {
            unsigned int w[kSjBatch];
            int c[kSjBatch];
            double v[kSjBatch], xv[kSjBatch];
#pragma unroll
            for (int k = 0; k < kSjBatch; ++k) {
                const int b = b0 + k * kSjBlock;
                w[k] = (b < ns) ? ld_stream(word + b) : 0u;
                c[k] = COL16 ? col0 + (int)(w[k] & 0xffffu) : ((b < ns) ? ld_stream(col + b) : 0);
            }
#pragma unroll
            for (int k = 0; k < kSjBatch; ++k) {
                const int b = b0 + k * kSjBlock;
                v[k] = (b < ns) ? ld_stream(val + b) : 0.0;
            }
#pragma unroll
            for (int k = 0; k < kSjBatch; ++k) xv[k] = ld_nc(a.x + c[k]);
#pragma unroll
            for (int k = 0; k < kSjBatch; ++k) {
                const int b = b0 + k * kSjBlock;
                if (b < ns) prod[COL16 ? (w[k] >> 16) : w[k]] = v[k] * xv[k];
            }
        }
        __syncthreads();
        if (tid < nrows) {
            const double* __restrict__ pt = prod + tid;
            double a0 = 0.0, a1 = 0.0;
            int d = 0;
            for (; d + 2 <= len; d += 2) {
                a0 += pt[sjd[d]];
                a1 += pt[sjd[d + 1]];
            }
            if (d < len) a0 += pt[sjd[d]];
            a.y[row] = dxr - (a0 + a1);
        }
        __syncthreads();
    }
}

// ---- K3 (persistent form): the whole Lanczos batch in ONE cooperative kernel -----------------------
// State per node i is one 32-byte sector  S_i = (z_i, u_i, u'_i, diag_i)  with
//     z = L u (un-normalised),  u = current Lanczos vector u_j,  u' = u_{j-1}.
// The next vector is never materialised on its own: every consumer evaluates
//     u_{j+1}[c] = k1 z_c + k2 u_c + k3 u'_c + k4,
//     k1 = 1/beta_j, k2 = -alpha_j/beta_j, k3 = -beta_j/beta_{j-1}, k4 = -(mean of the rest)
// from the gathered sector.  A gather of one double costs a 32-byte sector anyway, so packing the three
// operands of the recurrence into that sector makes the fused form free in L2 traffic, and one Lanczos
// step becomes ONE pass (gather + row sums) and ONE grid barrier (alpha_j, beta_j, sums).
// Phase j computes u_j, z_j = L u_j and the partial sums  p1 = u.z, p2 = sum z, p3 = u.u, p4 = sum u;
// after the barrier every CTA reduces the partials in the same fixed order, so all CTAs hold bitwise
// identical alpha_j = p1/p3, beta_j = sqrt(p3).
struct LzPersistState {
    int phase;        // phases completed so far (= number of alpha/beta entries valid)
    int cur;          // which sector buffer holds the input of the next phase
    double k1, k2, k3, k4;
    double beta_prev; // beta_{j-1} and sum(u_{j-1}) of the last completed phase j
    double usum_prev;
    unsigned int bar; // monotone barrier ticket counter
    unsigned int pad;
};

// One record per CTA and phase parity: the CTA's four partial sums plus a tag (= phase + 1) published with
// release semantics.  Polling every CTA's tag IS the grid barrier, and the data needed after the barrier
// arrives with it -- one L2 round trip instead of fence + atomic + poll + a separate reduction pass.
struct __align__(64) LzPartRec {
    double p[4];
    unsigned long long tag;
    unsigned long long pad[3];
};

struct LzPersistArgs {
    int n, ld, nphases, ncta;
    const int* rp;
    const int* col;
    const double* val;
    const int* row_start;   // [ncta + 1] contiguous row range per CTA, balanced by non-zeros
    double* sect[2];        // two sector buffers, 4 doubles per node
    double* basis;          // basis[j * ld + i] = u_j[i]
    double* alpha;
    double* beta;
    LzPartRec* recs;        // [2][ncta]
    LzPersistState* st;
    long long* timing;      // debug (MACB_PTIMING builds): [phase][cta][4] clock64 stamps
    // Asynchronous Rayleigh-Ritz: (alpha_j, beta_j) are streamed into host-mapped memory as they are produced and
    // the host raises *stop (host-mapped) once the Ritz pair has converged; CTA 0 samples it once per phase and
    // publishes its decision with its partial-sum record, so all CTAs leave after the same phase.
    double* ab_host;        // [2 * phase] = alpha, [2 * phase + 1] = beta   (may be nullptr)
    const int* stop;        // (may be nullptr)
};

__device__ __forceinline__ void ld_sector(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void st_sector(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.cg.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int ncta) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int old = atomicAdd(bar, 1u);
        unsigned int target = (old / ncta + 1u) * ncta;
        unsigned int cur;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(bar) : "memory");
        } while ((int)(cur - target) < 0);
        __threadfence();
    }
    __syncthreads();
}

// Predicated sector gather: lanes whose slot is past the row end, or whose edge weight is exactly zero
// (a candidate outside the current support -- most of the union pattern early in Frank-Wolfe), issue no
// memory request at all.
__device__ __forceinline__ void ld_sector_if(const double* p, bool pred, double& a, double& b, double& c) {
    double d;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %5, 0;\n\t"
        "mov.f64 %0, 0d0000000000000000;\n\tmov.f64 %1, 0d0000000000000000;\n\t"
        "mov.f64 %2, 0d0000000000000000;\n\tmov.f64 %3, 0d0000000000000000;\n\t"
        "@q ld.global.cg.v4.f64 {%0,%1,%2,%3}, [%4];\n\t}"
        : "=d"(a), "=d"(b), "=d"(c), "=d"(d)
        : "l"(p), "r"((int)pred));
}

constexpr int kPBlock = 1024;
constexpr int kPWarps = kPBlock / 32;

// Each sub-warp of W lanes walks its own row (W-strided segments of col/val).  The general fall-back: any row length, any size.
template <int W>
__global__ void __launch_bounds__(kPBlock, 1) k_lanczos_persist(LzPersistArgs a) {
    __shared__ double sm[4 * kPWarps];
    __shared__ double tot[4];
    __shared__ int stop_sm;
    constexpr int rpw = 32 / W;                    // rows per warp per pass
    const int lane = threadIdx.x & (W - 1);
    const int sub = (threadIdx.x & 31) / W;
    const int warp = threadIdx.x >> 5;
    const int r0 = a.row_start[blockIdx.x], r1 = a.row_start[blockIdx.x + 1];
    const int* __restrict__ col = a.col;
    const double* __restrict__ val = a.val;
    const int* __restrict__ rp = a.rp;

    int phase = a.st->phase;
    int cur = a.st->cur;
    double k1 = a.st->k1, k2 = a.st->k2, k3 = a.st->k3, k4 = a.st->k4;
    double beta_prev = a.st->beta_prev, usum_prev = a.st->usum_prev;   // beta_prev: 1/beta of the last completed step
    const double inv_n = 1.0 / (double)a.n;

    for (int it = 0; it < a.nphases; ++it, ++phase) {
        const double* __restrict__ S = a.sect[cur];
        double* __restrict__ D = a.sect[cur ^ 1];
        double* __restrict__ bj = a.basis + (size_t)phase * a.ld;
        double p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0;
        int stop_now = 0;
        if (blockIdx.x == 0 && threadIdx.x == 0 && a.stop)
            asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(stop_now) : "l"(a.stop));
#ifdef MACB_PTIMING
        long long t_start = clock64();
#endif
        for (int base = r0 + warp * rpw; base < r1; base += kPWarps * rpw) {
            const int row = base + sub;
            const bool valid = row < r1;
            double acc0 = 0.0, acc1 = 0.0;
            double oz = 0.0, ou = 0.0, oq = 0.0, od = 0.0;   // the row's own sector, requested before the gathers
            if (valid && lane == 0) ld_sector(S + 4 * (size_t)row, oz, ou, oq, od);
            if (valid) {
                const int s1 = rp[row + 1];
                for (int s = rp[row] + lane; s < s1; s += 4 * W) {
                    // four slots per lane in flight; out-of-row slots re-read slot s (same line) with weight 0
                    const bool v1 = s + W < s1, v2 = s + 2 * W < s1, v3 = s + 3 * W < s1;
                    const int i1 = v1 ? s + W : s, i2 = v2 ? s + 2 * W : s, i3 = v3 ? s + 3 * W : s;
                    const int c0 = ld_nc(col + s), c1 = ld_nc(col + i1), c2 = ld_nc(col + i2), c3 = ld_nc(col + i3);
                    const double w0 = ld_nc(val + s);
                    const double w1 = v1 ? ld_nc(val + i1) : 0.0, w2 = v2 ? ld_nc(val + i2) : 0.0,
                                 w3 = v3 ? ld_nc(val + i3) : 0.0;
                    double z0, u0, q0, z1, u1, q1, z2, u2, q2, z3, u3, q3;
                    ld_sector_if(S + 4 * (size_t)c0, w0 != 0.0, z0, u0, q0);
                    ld_sector_if(S + 4 * (size_t)c1, w1 != 0.0, z1, u1, q1);
                    ld_sector_if(S + 4 * (size_t)c2, w2 != 0.0, z2, u2, q2);
                    ld_sector_if(S + 4 * (size_t)c3, w3 != 0.0, z3, u3, q3);
                    // sum_s w_s (u_next[c_s] - k4) ; the constant k4 is folded back through diag below
                    acc0 = fma(w0, fma(k1, z0, fma(k2, u0, k3 * q0)), acc0);
                    acc1 = fma(w1, fma(k1, z1, fma(k2, u1, k3 * q1)), acc1);
                    acc0 = fma(w2, fma(k1, z2, fma(k2, u2, k3 * q2)), acc0);
                    acc1 = fma(w3, fma(k1, z3, fma(k2, u3, k3 * q3)), acc1);
                }
            }
            acc0 += acc1;
#pragma unroll
            for (int o = W / 2; o > 0; o >>= 1) acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
            if (valid && lane == 0) {
                const double z = oz, u = ou, q = oq, d = od;
                const double t = fma(k1, z, fma(k2, u, k3 * q));
                const double un = t + k4;                                    // u_phase[row]
                const double zn = fma(d, t, -acc0);                         // (L u_phase)[row]; L 1 = 0 cancels k4
                st_sector(D + 4 * (size_t)row, zn, un, u, d);
                __stcs(bj + row, un);   // streaming: the basis must not push the matrix out of L2
                p1 = fma(un, zn, p1);
                p2 += zn;
                p3 = fma(un, un, p3);
                p4 += un;
            }
        }
        // ---- CTA partials -> tagged record; poll all records (= grid barrier); same fixed-order sum everywhere
        double v[4] = {p1, p2, p3, p4};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
        if ((threadIdx.x & 31) == 0)
#pragma unroll
            for (int i = 0; i < 4; ++i) sm[i * kPWarps + warp] = v[i];
        __syncthreads();   // all row results of this CTA are written (sector buffer, basis) before the tag below
#ifdef MACB_PTIMING
        long long t_rows = clock64();
#endif
        LzPartRec* recs = a.recs + (size_t)(phase & 1) * a.ncta;
        const unsigned long long want = (unsigned long long)phase + 1ull;
        if (warp == 0) {
            double x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                x[i] = sm[i * kPWarps + (threadIdx.x & 31)];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x[i] += __shfl_xor_sync(0xffffffffu, x[i], o);
            }
            // arrive: publish the record, then one release-add on the shared counter; wait: poll that ONE counter
            // (a per-CTA flag poll would put ncta^2 readers on the L2), then fetch all records in one round trip.
            if (threadIdx.x == 0) {
                st_sector(recs[blockIdx.x].p, x[0], x[1], x[2], x[3]);
                if (blockIdx.x == 0) __stcg(&recs[0].pad[0], (unsigned long long)stop_now);
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&a.st->bar) : "memory");
                const unsigned int target = (unsigned int)want * (unsigned int)a.ncta;
                unsigned int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&a.st->bar) : "memory");
                } while ((int)(seen - target) < 0);
            }
            __syncwarp();
            double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
            for (int b = (threadIdx.x & 31); b < a.ncta; b += 32) {
                double q0, q1, q2, q3;
                ld_sector(recs[b].p, q0, q1, q2, q3);
                y0 += q0; y1 += q1; y2 += q2; y3 += q3;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                y0 += __shfl_xor_sync(0xffffffffu, y0, o);
                y1 += __shfl_xor_sync(0xffffffffu, y1, o);
                y2 += __shfl_xor_sync(0xffffffffu, y2, o);
                y3 += __shfl_xor_sync(0xffffffffu, y3, o);
            }
            if (threadIdx.x == 0) {
                tot[0] = y0; tot[1] = y1; tot[2] = y2; tot[3] = y3;
                stop_sm = (int)__ldcg(&recs[0].pad[0]);
            }
        }
        __syncthreads();
#ifdef MACB_PTIMING
        long long t_bar = clock64();
#endif
        const double P1 = tot[0], P2 = tot[1], P3 = tot[2], P4 = tot[3];
        const int stop_all = stop_sm;
        const LzCoef cf = lz_coefficients(P1, P2, P3, P4, (phase > 0) ? beta_prev : 0.0, usum_prev, inv_n);
        const double alpha = cf.alpha, beta = cf.beta;
        const double nk1 = cf.k1, nk2 = cf.k2, nk3 = cf.k3, nk4 = cf.k4;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            a.alpha[phase] = alpha;
            a.beta[phase] = beta;
            if (a.ab_host)   // one 16-byte posted write: the host sees alpha and beta together
                asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(a.ab_host + 2 * (size_t)phase), "d"(alpha), "d"(beta) : "memory");
        }
#ifdef MACB_PTIMING
        if (threadIdx.x == 0 && a.timing && it < 64) {
            long long* t = a.timing + ((size_t)it * a.ncta + blockIdx.x) * 4;
            t[0] = t_start; t[1] = t_rows; t[2] = t_bar; t[3] = clock64();
        }
#endif
        k1 = nk1; k2 = nk2; k3 = nk3; k4 = nk4;
        beta_prev = cf.binv;   // (holds 1/beta of the completed step)
        usum_prev = P4;
        cur ^= 1;
        if (stop_all) {
            ++phase;
            break;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.st->phase = phase;
        if (a.stop) const_cast<int*>(a.stop)[1] = phase;   // host-mapped: read after the stream synchronise, no extra copy
        a.st->cur = cur;
        a.st->k1 = k1; a.st->k2 = k2; a.st->k3 = k3; a.st->k4 = k4;
        a.st->beta_prev = beta_prev;
        a.st->usum_prev = usum_prev;
    }
}

// ---- K3, slot-parallel form (default) -------------------------------------------------------------------
// Same state, recurrence, barrier and outputs as k_lanczos_persist, different mapping of work inside a CTA:
//   pass 1  one thread per SLOT of the CTA's contiguous slot range, four slots in flight per thread, fully
//           coalesced (col, val) reads; the product  w_s * (u_next[c_s] - k4)  goes to shared memory;
//   pass 2  one thread per ROW sums its segment from shared memory (no shuffles, no global latency: the row's own
//           sector was requested before pass 1) and writes the new sector, the basis entry and the partial sums.
// The gather loop is then exactly the loop tools/micro/gather_mix.cu measures at 0.9 sector/clk/SM -- the
// divergent-gather ceiling of the LSU/L1TEX path on B200 -- instead of stalling on a per-row reduction after
// every four gathers.  A CTA's range is cut into chunks of <= kPBlock rows and <= cap slots (shared memory).
struct LzChunkArgs {
    const int* chunk_ptr;   // [ncta + 1] chunks of CTA b are chunk_ptr[b] .. chunk_ptr[b+1]
    const int* chunk_row;   // [nchunks + 1] first row of every chunk (rows of chunk q: chunk_row[q] .. chunk_row[q+1])
    int cache_cols;         // every CTA has one chunk and 12 bytes/slot fit in shared memory: column indices (with an
                            // "inactive" bit for zero weights) stay in shared memory for the whole launch, so a gather
                            // address never waits for a global load
    int prod_cap;           // slots reserved for the product buffer (the column cache follows it)
};

__global__ void __launch_bounds__(kPBlock, 1) k_lanczos_slots(LzPersistArgs a, LzChunkArgs ch) {
    extern __shared__ double prod[];
    __shared__ double sm[4 * kPWarps];
    __shared__ double tot[4];
    __shared__ int stop_sm;
    const int warp = threadIdx.x >> 5;
    const int* __restrict__ col = a.col;
    const double* __restrict__ val = a.val;
    const int* __restrict__ rp = a.rp;
    const int q0 = ch.chunk_ptr[blockIdx.x], q1 = ch.chunk_ptr[blockIdx.x + 1];
    int* __restrict__ scol = reinterpret_cast<int*>(prod + ch.prod_cap);
    if (ch.cache_cols && q1 > q0) {
        const int sa = rp[ch.chunk_row[q0]], sb = rp[ch.chunk_row[q0 + 1]];
        for (int i = sa + (int)threadIdx.x; i < sb; i += kPBlock)
            scol[i - sa] = ld_nc(col + i) | ((ld_nc(val + i) == 0.0) ? (int)0x80000000 : 0);
        __syncthreads();
    }

    int phase = a.st->phase;
    int cur = a.st->cur;
    double k1 = a.st->k1, k2 = a.st->k2, k3 = a.st->k3, k4 = a.st->k4;
    double beta_prev = a.st->beta_prev, usum_prev = a.st->usum_prev;   // beta_prev: 1/beta of the last completed step
    const double inv_n = 1.0 / (double)a.n;

    for (int it = 0; it < a.nphases; ++it, ++phase) {
        const double* __restrict__ S = a.sect[cur];
        double* __restrict__ D = a.sect[cur ^ 1];
        double* __restrict__ bj = a.basis + (size_t)phase * a.ld;
        double p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0;
        int stop_now = 0;
        if (blockIdx.x == 0 && threadIdx.x == 0 && a.stop)
            asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(stop_now) : "l"(a.stop));
#ifdef MACB_PTIMING
        long long t_start = clock64(), t_p1 = 0;
#endif
        if (ch.cache_cols) {
            // single chunk, column indices resident in shared memory
            const int ra = ch.chunk_row[q0], rb = ch.chunk_row[q0 + 1];
            const int row = ra + (int)threadIdx.x;
            const bool has_row = (q1 > q0) && row < rb;
            double oz = 0.0, ou = 0.0, oq = 0.0, od = 0.0;
            if (has_row) ld_sector(S + 4 * (size_t)row, oz, ou, oq, od);
            const int sa = (q1 > q0) ? rp[ra] : 0, sb = (q1 > q0) ? rp[rb] : 0;
            const int ns = sb - sa;
            for (int j0 = (int)threadIdx.x; j0 < ns; j0 += 4 * kPBlock) {
                const int j1 = j0 + kPBlock, j2 = j0 + 2 * kPBlock, j3 = j0 + 3 * kPBlock;
                const bool v1 = j1 < ns, v2 = j2 < ns, v3 = j3 < ns;
                const int c0 = scol[j0], c1 = v1 ? scol[j1] : (int)0x80000000, c2 = v2 ? scol[j2] : (int)0x80000000,
                          c3 = v3 ? scol[j3] : (int)0x80000000;
                double z0, u0, g0, z1, u1, g1, z2, u2, g2, z3, u3, g3;
                ld_sector_if(S + 4 * (size_t)(c0 & 0x7fffffff), c0 >= 0, z0, u0, g0);
                ld_sector_if(S + 4 * (size_t)(c1 & 0x7fffffff), c1 >= 0, z1, u1, g1);
                ld_sector_if(S + 4 * (size_t)(c2 & 0x7fffffff), c2 >= 0, z2, u2, g2);
                ld_sector_if(S + 4 * (size_t)(c3 & 0x7fffffff), c3 >= 0, z3, u3, g3);
                const double w0 = (c0 >= 0) ? ld_nc(val + sa + j0) : 0.0, w1 = (c1 >= 0) ? ld_nc(val + sa + j1) : 0.0,
                             w2 = (c2 >= 0) ? ld_nc(val + sa + j2) : 0.0, w3 = (c3 >= 0) ? ld_nc(val + sa + j3) : 0.0;
                prod[j0] = w0 * fma(k1, z0, fma(k2, u0, k3 * g0));
                if (v1) prod[j1] = w1 * fma(k1, z1, fma(k2, u1, k3 * g1));
                if (v2) prod[j2] = w2 * fma(k1, z2, fma(k2, u2, k3 * g2));
                if (v3) prod[j3] = w3 * fma(k1, z3, fma(k2, u3, k3 * g3));
            }
            __syncthreads();
#ifdef MACB_PTIMING
            t_p1 = clock64();
#endif
            if (has_row) {
                const int s0 = rp[row] - sa, s1 = rp[row + 1] - sa;
                double acc0 = 0.0, acc1 = 0.0;
                int i = s0;
                for (; i + 1 < s1; i += 2) {
                    acc0 += prod[i];
                    acc1 += prod[i + 1];
                }
                if (i < s1) acc0 += prod[i];
                const double t = fma(k1, oz, fma(k2, ou, k3 * oq));
                const double un = t + k4;
                const double zn = fma(od, t, -(acc0 + acc1));
                st_sector(D + 4 * (size_t)row, zn, un, ou, od);
                __stcs(bj + row, un);   // streaming: the basis must not push the matrix out of L2
                p1 = fma(un, zn, p1);
                p2 += zn;
                p3 = fma(un, un, p3);
                p4 += un;
            }
        } else
        for (int q = q0; q < q1; ++q) {
            const int ra = ch.chunk_row[q], rb = ch.chunk_row[q + 1];
            const int row = ra + (int)threadIdx.x;
            const bool has_row = row < rb;
            double oz = 0.0, ou = 0.0, oq = 0.0, od = 0.0;
            if (has_row) ld_sector(S + 4 * (size_t)row, oz, ou, oq, od);
            const int sa = rp[ra], sb = rp[rb];
            for (int i0 = sa + (int)threadIdx.x; i0 < sb; i0 += 4 * kPBlock) {
                const int i1 = i0 + kPBlock, i2 = i0 + 2 * kPBlock, i3 = i0 + 3 * kPBlock;
                const bool v1 = i1 < sb, v2 = i2 < sb, v3 = i3 < sb;
                const int c0 = ld_nc(col + i0), c1 = v1 ? ld_nc(col + i1) : 0, c2 = v2 ? ld_nc(col + i2) : 0,
                          c3 = v3 ? ld_nc(col + i3) : 0;
                const double w0 = ld_nc(val + i0), w1 = v1 ? ld_nc(val + i1) : 0.0, w2 = v2 ? ld_nc(val + i2) : 0.0,
                             w3 = v3 ? ld_nc(val + i3) : 0.0;
                double z0, u0, g0, z1, u1, g1, z2, u2, g2, z3, u3, g3;
                ld_sector_if(S + 4 * (size_t)c0, w0 != 0.0, z0, u0, g0);
                ld_sector_if(S + 4 * (size_t)c1, w1 != 0.0, z1, u1, g1);
                ld_sector_if(S + 4 * (size_t)c2, w2 != 0.0, z2, u2, g2);
                ld_sector_if(S + 4 * (size_t)c3, w3 != 0.0, z3, u3, g3);
                prod[i0 - sa] = w0 * fma(k1, z0, fma(k2, u0, k3 * g0));
                if (v1) prod[i1 - sa] = w1 * fma(k1, z1, fma(k2, u1, k3 * g1));
                if (v2) prod[i2 - sa] = w2 * fma(k1, z2, fma(k2, u2, k3 * g2));
                if (v3) prod[i3 - sa] = w3 * fma(k1, z3, fma(k2, u3, k3 * g3));
            }
            __syncthreads();
            if (has_row) {
                const int s0 = rp[row] - sa, s1 = rp[row + 1] - sa;
                double acc0 = 0.0, acc1 = 0.0;
                int i = s0;
                for (; i + 1 < s1; i += 2) {
                    acc0 += prod[i];
                    acc1 += prod[i + 1];
                }
                if (i < s1) acc0 += prod[i];
                const double t = fma(k1, oz, fma(k2, ou, k3 * oq));
                const double un = t + k4;                       // u_phase[row]
                const double zn = fma(od, t, -(acc0 + acc1));   // (L u_phase)[row]; L 1 = 0 cancels k4
                st_sector(D + 4 * (size_t)row, zn, un, ou, od);
                __stcs(bj + row, un);   // streaming: the basis must not push the matrix out of L2
                p1 = fma(un, zn, p1);
                p2 += zn;
                p3 = fma(un, un, p3);
                p4 += un;
            }
            __syncthreads();
        }
        // ---- identical to k_lanczos_persist from here: tagged record, counter barrier, fixed-order reduction
        double v[4] = {p1, p2, p3, p4};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
        if ((threadIdx.x & 31) == 0)
#pragma unroll
            for (int i = 0; i < 4; ++i) sm[i * kPWarps + warp] = v[i];
        __syncthreads();
#ifdef MACB_PTIMING
        long long t_rows = clock64();
#endif
        LzPartRec* recs = a.recs + (size_t)(phase & 1) * a.ncta;
        const unsigned long long want = (unsigned long long)phase + 1ull;
        if (warp == 0) {
            double x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                x[i] = sm[i * kPWarps + (threadIdx.x & 31)];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) x[i] += __shfl_xor_sync(0xffffffffu, x[i], o);
            }
            if (threadIdx.x == 0) {
                st_sector(recs[blockIdx.x].p, x[0], x[1], x[2], x[3]);
                if (blockIdx.x == 0) __stcg(&recs[0].pad[0], (unsigned long long)stop_now);
                asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(&a.st->bar) : "memory");
                const unsigned int target = (unsigned int)want * (unsigned int)a.ncta;
                unsigned int seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&a.st->bar) : "memory");
                } while ((int)(seen - target) < 0);
            }
            __syncwarp();
            double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
            for (int b = (threadIdx.x & 31); b < a.ncta; b += 32) {
                double r0, r1, r2, r3;
                ld_sector(recs[b].p, r0, r1, r2, r3);
                y0 += r0; y1 += r1; y2 += r2; y3 += r3;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                y0 += __shfl_xor_sync(0xffffffffu, y0, o);
                y1 += __shfl_xor_sync(0xffffffffu, y1, o);
                y2 += __shfl_xor_sync(0xffffffffu, y2, o);
                y3 += __shfl_xor_sync(0xffffffffu, y3, o);
            }
            if (threadIdx.x == 0) {
                tot[0] = y0; tot[1] = y1; tot[2] = y2; tot[3] = y3;
                stop_sm = (int)__ldcg(&recs[0].pad[0]);
            }
        }
        __syncthreads();
#ifdef MACB_PTIMING
        long long t_bar = clock64();
#endif
        const double P1 = tot[0], P2 = tot[1], P3 = tot[2], P4 = tot[3];
        const int stop_all = stop_sm;
        const LzCoef cf = lz_coefficients(P1, P2, P3, P4, (phase > 0) ? beta_prev : 0.0, usum_prev, inv_n);
        const double alpha = cf.alpha, beta = cf.beta;
        const double nk1 = cf.k1, nk2 = cf.k2, nk3 = cf.k3, nk4 = cf.k4;
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            a.alpha[phase] = alpha;
            a.beta[phase] = beta;
            if (a.ab_host)
                asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(a.ab_host + 2 * (size_t)phase), "d"(alpha), "d"(beta) : "memory");
        }
#ifdef MACB_PTIMING
        if (threadIdx.x == 0 && a.timing && it < 64) {
            long long* t = a.timing + ((size_t)it * a.ncta + blockIdx.x) * 4;
            t[0] = t_start; t[1] = t_rows; t[2] = t_bar; t[3] = clock64();
            a.timing[(size_t)64 * a.ncta * 4 + (size_t)it * a.ncta + blockIdx.x] = t_p1;
        }
#endif
        k1 = nk1; k2 = nk2; k3 = nk3; k4 = nk4;
        beta_prev = cf.binv;   // (holds 1/beta of the completed step)
        usum_prev = P4;
        cur ^= 1;
        if (stop_all) {
            ++phase;
            break;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.st->phase = phase;
        if (a.stop) const_cast<int*>(a.stop)[1] = phase;   // host-mapped: read after the stream synchronise, no extra copy
        a.st->cur = cur;
        a.st->k1 = k1; a.st->k2 = k2; a.st->k3 = k3; a.st->k4 = k4;
        a.st->beta_prev = beta_prev;
        a.st->usum_prev = usum_prev;
    }
}

// ---- layout of the pipelined Lanczos kernel (k_lanczos_pipe, below): host side in build_slice_layout (api.cu) ----------
// History of the product buffer, all measured on B200 at the headline size: CSR order (4-way bank conflicts in the row sums)
// -> jagged diagonals (round 1: consecutive threads read consecutive words, uniform trip counts, but a table of diagonal
// starts between the loads) -> 32-row slices with an odd lane stride (this round: addresses known up front, no predication).
constexpr int kLzSlice = 33;      // lane stride of a slice (doubles): odd, so that the entry index moves the shared-memory bank
constexpr int kLzSliceTab = 64;   // ints per CTA in the slice table: (base, length) of up to 32 slices
struct LzJdsArgs {
    const int* row_start;   // [ncta + 1] rows of CTA b (<= kPBlock of them), in the ENGINE's node numbering
    const int* jcol;        // [nnz] engine column | product position << 17 of every slot, column order inside a CTA's slot range
    const double* jval;     // [nnz] weight of every slot, same order (k_assemble_jds)
    const int* jw;          // [ncta * kLzSliceTab] slice table: (first position, entries per row) of warp w's 32 rows
    int pos_cap;            // product positions reserved in shared memory (doubles); the column cache and the slice table follow
    int slot_cap;           // slots reserved for the column cache (ints)
    double* xrec;           // [2][ncta][ncta][4] inboxes of the all-to-all barrier, NaN = empty (k_lz_persist_init)
    const double* diag;     // [n] weighted degrees, caller numbering 
    const int* perm;        // [n] engine -> caller numbering
};
// Engine numbering: inside every CTA's row range the rows are renumbered by decreasing length (perm[new] = old), so
// that thread t owns engine row ra + t and its sector / basis accesses stay coalesced.  k_lz_persist_init permutes
// the start vector in, k_ritz permutes the Ritz vector out; nothing outside the Lanczos engine sees the numbering.

__global__ void __launch_bounds__(kBlock) k_assemble_jds(int64_t nnz, const int* __restrict__ jeid,
                                                         const double* __restrict__ ew, double* __restrict__ jval) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < nnz; s += (int64_t)gridDim.x * blockDim.x)
        jval[s] = ew[ld_stream(jeid + s)];
}

// Sum of four values over the 32 lanes of a warp with 6 double shuffles instead of 20: two halving steps in which
// every lane sends away half of the values it still holds, then three plain butterfly steps on the one value left.
// Every lane l ends with the warp total of value (l >> 3) & 3 ... precisely: bit 4 of the lane selects {0,1} vs {2,3},
// bit 3 selects the even or odd member.  The shuffle network (SHFL) shares the MIO path with the gathers, so fewer
// shuffles is fewer stalled cycles at the end of every step.
__device__ __forceinline__ double warp_sum4(double v0, double v1, double v2, double v3, int lane) {
    const bool hi = lane & 16;
    // lanes with bit 4 set keep (v2, v3) and give away (v0, v1); the others the reverse
    double s0 = hi ? v0 : v2, s1 = hi ? v1 : v3;
    double k0 = hi ? v2 : v0, k1 = hi ? v3 : v1;
    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    const bool mid = lane & 8;
    double s = mid ? k0 : k1, k = mid ? k1 : k0;
    k += __shfl_xor_sync(0xffffffffu, s, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;   // lane l holds the total of value 2 * (l >> 4) + ((l >> 3) & 1)
}

// ---- K3, pipelined form (default): the reduction leaves the critical path ---------------------------------------
// Its predecessor (round 1's k_lanczos_vec: same layout, plain Lanczos) spent a quarter of every step in the grid-wide exchange of the four partial sums (three dependent L2 round
// trips) plus the wait for the slowest CTA, because the coefficients alpha_j, beta_j of step j depend on z_j = L u_j, the
// result of that very step's SpMV.  Here the recurrence is rearranged (Ghysels/Vanroose-style pipelining) so that the SpMV
// and the reduction of a step are independent of each other:
//     state per node:  u_j, u_{j-1},  z_j = L' u_j,  z_{j-1}            (L' = P L P - sigma I on 1-perp)
//     step j:   sums (u_j.z_j, sum z_j, u_j.u_j, sum u_j) -> exchange          ... in flight during ...
//               q_j = L' z_j                                                   ... this SpMV (gathers of z_j)
//               u_{j+1} = k1 z_j + k2 u_j + k3 u_{j-1} + k4
//               z_{j+1} = k1 q_j + k2 z_j + k3 z_{j-1} - sigma k4       (L' 1 = -sigma 1)
// i.e. z_{j+1} = L' u_{j+1} is obtained by applying the three-term recurrence to the z's instead of by a product of its own.
// The records of step j are pushed at the very START of the step (they depend only on what the update of step j-1 left in
// registers) and are polled after the row sums, a whole SpMV later.
//
// Why the shift: the rounding errors d_j = z_j - L' u_j obey the same three-term recurrence as the Lanczos polynomials
// evaluated at the shift, d_{j+1} = -((alpha_j - sigma) d_j + beta_j d_{j-1}) / beta_{j+1}.  At sigma = 0 that point lies outside
// the spectrum of P L P on 1-perp and d_j grows like 1.25^j (measured: relative drift 1e-2 after 100 steps); with sigma inside
// the support of the spectrum -- the mean diagonal entry, trace(L)/n -- the polynomials stay O(1): measured drift 5e-15 after
// 260 steps at the headline size and 3e-14 after 4200 steps on city10000, Ritz values equal to plain Lanczos to 1e-15.
// alpha_j is published with the shift added back, so T_k, the Rayleigh-Ritz step and the stopping rule are unchanged.
//
// No fences, no poison stores: "has the producer written this yet?" is answered by a one-bit GENERATION TAG in the least
// significant mantissa bit of every published double.  The vector buffers alternate, so the buffer gathered in phase p was
// last written two phases ago: tag(p) = (p >> 1) & 1 differs from the stale content's tag, and L2 (the point of coherence: all
// these accesses bypass L1) never shows a reader an older value than one it has already seen.  The owner keeps the TAGGED
// value in its own registers, so every CTA multiplies with the same z_j; the perturbation is one ulp of the gathered operand,
// the size of the rounding error of the product it enters.  The records carry the same tag in the first of their four doubles
// (a 32-byte sector is written and read as one transaction).  The predecessor needed a release fence per step (2 000 cycles
// with a thousand stores in flight, measured) plus 8 + 32 bytes of poison per node and inbox slot to get the same guarantee.
__device__ __forceinline__ double lz_tagged(double v, int tag) {
    return __longlong_as_double((__double_as_longlong(v) & ~1ll) | (long long)tag);
}
__device__ __forceinline__ bool lz_tag_ok(double v, int tag) { return (int)(__double_as_longlong(v) & 1ll) == tag; }
__device__ __forceinline__ double ld_f64_relaxed_if(const double* p, bool pred, double other) {
    double v;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\tmov.f64 %0, %3;\n\t@q ld.relaxed.gpu.global.f64 %0, [%1];\n\t}"
                 : "=d"(v) : "l"(p), "r"((int)pred), "d"(other) : "memory");
    return v;
}

// One word per slot in the shared-memory column cache: engine column (17 bits; all ones = the slot's weight is zero in this
// launch, nothing to gather or store) | position of its product << 17 (15 bits).
constexpr int kLzColMask = 0x1ffff;
template <int VB>
__device__ __forceinline__ void lz_gather_products_tagged(const double* __restrict__ U, const int* __restrict__ scol,
                                                          const double* __restrict__ jval, int ns, int tid, int tag,
                                                          double* __restrict__ prod, int* give_up) {
    const double ok_zero = __longlong_as_double((long long)tag);   // what an unissued gather "returns": +0 with a valid tag
    for (int b0 = tid; b0 < ns; b0 += VB * kPBlock) {
        int c[VB];
        double v[VB], wq[VB];
#pragma unroll
        for (int q = 0; q < VB; ++q) {
            const int b = b0 + q * kPBlock;
            c[q] = (b < ns) ? scol[b] : kLzColMask;
        }
#pragma unroll
        for (int q = 0; q < VB; ++q) v[q] = ld_f64_relaxed_if(U + (c[q] & kLzColMask), (c[q] & kLzColMask) != kLzColMask, ok_zero);
        // the weights of the whole batch are requested up front as well: loaded one by one next to their use they form a
        // chain of VB dependent L2 latencies per batch
#pragma unroll
        for (int q = 0; q < VB; ++q) wq[q] = ((c[q] & kLzColMask) != kLzColMask) ? ld_stream(jval + b0 + q * kPBlock) : 0.0;
#pragma unroll
        for (int q = 0; q < VB; ++q) {
            if (!lz_tag_ok(v[q], tag)) {   // producer has not written yet: gather again (bounded: never hang the device)
                unsigned int tries = 0;
                do {
                    __nanosleep(100);   // the producer is still in its row sums: do not fill the memory pipe with polls
                    v[q] = ld_f64_relaxed_if(U + (c[q] & kLzColMask), true, ok_zero);
                } while (!lz_tag_ok(v[q], tag) && ++tries < (1u << 16));
                if (!lz_tag_ok(v[q], tag)) atomicExch(give_up, 1);
            }
            // inactive slots store nothing: their positions keep the zero written when the launch began
            if ((c[q] & kLzColMask) != kLzColMask) prod[(unsigned int)c[q] >> 17] = wq[q] * v[q];
        }
    }
}

// sum of entries d0 <= d < d1 of one row of a slice: p points at entry 0 of the lane's row, entries are kLzSlice doubles apart.
// All addresses are known up front (no table of diagonal starts between the loads), all lanes of the warp run the same trip
// count (the slice is padded with zeros to its longest row), nothing is predicated.
__device__ __forceinline__ double lz_slice_sum(const double* __restrict__ p, int d0, int d1) {
    const double* __restrict__ q = p + d0 * kLzSlice;
    int n = d1 - d0;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0, a5 = 0.0, a6 = 0.0, a7 = 0.0;
    for (; n >= 8; n -= 8, q += 8 * kLzSlice) {
        a0 += q[0]; a1 += q[kLzSlice]; a2 += q[2 * kLzSlice]; a3 += q[3 * kLzSlice];
        a4 += q[4 * kLzSlice]; a5 += q[5 * kLzSlice]; a6 += q[6 * kLzSlice]; a7 += q[7 * kLzSlice];
    }
    if (n & 4) {
        a0 += q[0]; a1 += q[kLzSlice]; a2 += q[2 * kLzSlice]; a3 += q[3 * kLzSlice];
        q += 4 * kLzSlice;
    }
    if (n & 2) {
        a4 += q[0]; a5 += q[kLzSlice];
        q += 2 * kLzSlice;
    }
    if (n & 1) a6 += q[0];
    return ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

struct RrOut {
    int status;      // 0 running; 1 estimate below target; 2 cycle exhausted (k_limit); 3 invariant subspace (breakdown);
                     // -1 non-finite coefficient; -2 eigenvector recurrence failed; -3 timed out waiting for coefficients
    int k;           // order of the T_k the decision (and coef) refers to
    int checks;
    int phases;      // phases the solver completed (written by solver CTA 0 at exit)
    double theta, est, target;
    long long cyc_wait, cyc_compute;   // diagnostics: cycles spent waiting for coefficients / computing
    int lag;                           // coefficients the solver had published beyond `k` when the decision fell
    int rounds;                        // multisection rounds, all checks
    long long cyc_stage[6];            // fetch, bounds, warm bracket, multisection, twisted factorisation + sweeps, rest
};

struct RrArgs {
    const double* alpha;   // [cap + 1] published by the solver, NaN = not yet
    const double* beta;    // [cap + 2]
    double* a;             // private copies / derived arrays, [cap + 2] each
    double* b;
    double* b2;
    double* binv;
    double* s;
    double* dp;            // eigenvector recurrences from the top / from the bottom, [cap + 2] each (global fall-back of the
    double* dm;            // shared-memory copies)
    double* coef;          // out [cap + 1]: s_t / beta_t
    RrOut* out;
    int* dev_stop;
    const LzScalars* sc;   // sc->lnorm
    double tol;
    int n, k_limit, check_div, enabled;
    int smem_doubles;      // dynamic shared memory of the launch, in doubles (the Rayleigh-Ritz CTA keeps T_k there)
};

struct LzPipeArgs {
    const LzScalars* sc;   // sc->shift = trace(L)/n (k_assemble)
    double* zprev;         // [n] z_{j-1} of the CTA's rows across launches (engine numbering)
    int* dev_stop;         // device-resident stop flag raised by the Rayleigh-Ritz CTA (may be nullptr)
};

// Start of a cycle: z_0 = L' u_0 comes from one plain SpMV (y = L u_0, caller numbering) so that the persistent kernel has no
// special first step.  Buffer 0 <- z_0 tagged 0, buffer 1 and the inboxes <- NaN with the tag bit SET (never valid for the
// phases 0 and 1 that read them first), basis row 0 <- u_0.
__global__ void __launch_bounds__(kBlock) k_lz_pipe_init(int n, const double* __restrict__ src, const double* __restrict__ lsrc,
                                                         const int* __restrict__ perm, const LzScalars* sc, double* __restrict__ z0,
                                                         double* __restrict__ z1, double* __restrict__ basis0, double* __restrict__ xrec,
                                                         int64_t nxrec, LzPersistState* st, double* __restrict__ alpha,
                                                         double* __restrict__ beta, int ncoef, RrOut* rr_out, int* dev_stop) {
    const double nan1 = __longlong_as_double(0x7ff8000000000001ll);
    const double sigma = sc->shift;
    // coefficients not yet produced read as NaN (the Rayleigh-Ritz CTA waits for every entry on its own)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncoef; i += gridDim.x * blockDim.x) {
        alpha[i] = nan1;
        beta[i] = nan1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && rr_out) {
        rr_out->status = 0; rr_out->k = 0; rr_out->checks = 0; rr_out->phases = 0;
        rr_out->theta = 0.0; rr_out->est = 0.0; rr_out->target = 0.0;
        rr_out->cyc_wait = 0; rr_out->cyc_compute = 0; rr_out->lag = 0; rr_out->rounds = 0;
        *dev_stop = 0;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = perm[i];
        const double u = src[c];
        z0[i] = lz_tagged(fma(-sigma, u, lsrc[c]), 0);
        z1[i] = nan1;
        basis0[i] = u;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nxrec; i += (int64_t)gridDim.x * blockDim.x) xrec[i] = nan1;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->phase = 0;
        st->cur = 0;
        st->k1 = 0.0; st->k2 = 1.0; st->k3 = 0.0; st->k4 = 0.0;
        st->beta_prev = 0.0;
        st->usum_prev = 0.0;
        st->bar = 0u;
    }
}

// ---- On-device Rayleigh-Ritz: the stop decision of the Lanczos iteration (replaces the 4x4 `eigh` + residual test of the
// reference's loop, nx:239-245, and the host thread that used to poll the coefficients over PCIe) ---------------------------------
// One extra CTA of the cooperative Lanczos launch (block index == number of solver CTAs).  It follows the coefficients
// (alpha_j, beta_j) that solver CTA 0 publishes in device memory, and at check points that depend on the coefficient sequence
// only (so a solve is a pure function of its input, whatever the timing) computes the smallest eigenpair (theta, s) of T_k:
//   theta   1024-way multisection on the Sturm count (every thread one shift; the bracket shrinks 1025x per round), started
//           from a geometric bracket below the previous theta (Ritz values decrease monotonically with k);
//   s       forward pivot recurrence (stable for the smallest eigenvalue: T_j - theta I is positive semi-definite for every
//           leading block), one thread;
//   est     |beta_k s_{k-1}| = 2-norm of the residual of the Ritz pair; stop when 0.88 est sqrt(n) < tol ||L||_inf (the
//           reference's test ||r||_1 / ||L||_inf < tol with ||r||_1 ~ 0.80 sqrt(n) ||r||_2 for a Gaussian-like residual and
//           10 % margin; the TRUE residual is tested after the kernel in any case).
// On the decision it raises a device-resident flag that solver CTA 0 reads once per phase, and leaves k, theta and the Ritz
// coefficients s_t / beta_t for k_ritz.  The host sees nothing of this until it reads the result of the whole FW iteration.
__device__ __forceinline__ double ld_relaxed_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// block-wide (1024 threads) max of a value; result on every thread
__device__ __forceinline__ double rr_block_max(double v, double* red /*[32]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmax(r, red[w]);
    return r;
}

__device__ __forceinline__ double rr_block_sum(double v, double* red /*[32]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += red[w];
    return r;
}

// true iff T_k has an eigenvalue below x (Sturm sequence, early exit on the first negative pivot)
// 1/x to about one ulp: hardware seed (MUFU.RCP64H, ~20 bits) and two Newton steps.  An IEEE double division is a ~40-instruction
// sequence with ~250 cycles of latency on this chip, and the Sturm recurrence below is one long chain of them: the check of a
// T_190 took 180 us and the solver overshot its stopping point by 60 % (measured); the recurrence is backward stable with
// respect to such last-bit errors (they are relative perturbations of the matrix entries of the same size).
// Eigenvector of T_k for theta without a single division: the three-term recurrence run from the top (f) and from the bottom
// (g), one thread each, joined at the index r of the largest |f| -- the vector of the twisted factorisation at that r, obtained
// through the minors instead of the pivots: ~25 cycles per row instead of ~100 (a pivot costs a reciprocal).  The recurrence from
// the top is accurate until the entries start to decay (the decaying solution is the recessive one), which is after their
// maximum; the one from the bottom is accurate all the way up to there (it follows the growing solution).
__device__ __forceinline__ void rr_three_term_down(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ binv,
                                                   int k, double theta, double* __restrict__ f) {
    double f0 = 1.0, f1 = (k > 1) ? -(a[0] - theta) * binv[1] : 0.0;
    f[0] = f0;
    if (k > 1) f[1] = f1;
    for (int i = 1; i + 1 < k; ++i) {
        const double fn = -fma(a[i] - theta, f1, b[i] * f0) * binv[i + 1];
        f[i + 1] = fn;
        f0 = f1;
        f1 = fn;
    }
}
__device__ __forceinline__ void rr_three_term_up(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ binv,
                                                 int k, double theta, double* __restrict__ g) {
    double g0 = 1.0, g1 = (k > 1) ? -(a[k - 1] - theta) * binv[k - 1] : 0.0;   // g0 = g_{k-1}, g1 = g_{k-2}
    g[k - 1] = g0;
    if (k > 1) g[k - 2] = g1;
    for (int i = k - 2; i >= 1; --i) {
        const double gn = -fma(a[i] - theta, g1, b[i + 1] * g0) * binv[i];
        g[i - 1] = gn;
        g0 = g1;
        g1 = gn;
        if (fabs(gn) > 1e200) {   // grows towards the top: rescale what has been written so far (rare)
            for (int t = i - 1; t < k; ++t) g[t] *= 1e-200;
            g0 *= 1e-200;
            g1 *= 1e-200;
        }
    }
}

// The same two recurrences with every array in SHARED memory (32-bit shared addresses), four rows per trip, the entries of the
// next trip requested before the current one is worked on: the chain is then the two dependent floating-point operations per
// row.  (Through generic pointers the compiler keeps the loads inside the chain: 240 cycles per row, measured.)
__device__ __forceinline__ double rr_lds(unsigned int addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void rr_sts(unsigned int addr, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory"); }

__device__ __forceinline__ void rr_three_term_down_smem(unsigned int a_sh, unsigned int b_sh, unsigned int binv_sh, unsigned int f_sh, int k,
                                                        double theta) {
    double f0 = 1.0;
    rr_sts(f_sh, 1.0);
    if (k < 2) return;
    double f1 = -(rr_lds(a_sh) - theta) * rr_lds(binv_sh + 8u);
    rr_sts(f_sh + 8u, f1);
    double an[4], bn[4], vn[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // rows 1..4 (the arrays reach at least k + 8: reads past the end are harmless)
        an[j] = rr_lds(a_sh + 8u * (1 + j));
        bn[j] = rr_lds(b_sh + 8u * (1 + j));
        vn[j] = rr_lds(binv_sh + 8u * (2 + j));
    }
    for (int i = 1; i + 1 < k; i += 4) {
        double ac[4], bc[4], vc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { ac[j] = an[j]; bc[j] = bn[j]; vc[j] = vn[j]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            an[j] = rr_lds(a_sh + 8u * (unsigned int)(i + 4 + j));
            bn[j] = rr_lds(b_sh + 8u * (unsigned int)(i + 4 + j));
            vn[j] = rr_lds(binv_sh + 8u * (unsigned int)(i + 5 + j));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i + j + 1 < k) {
                // one fused multiply-add on the dependency chain: the two products below do not depend on f1
                const double fn = fma((theta - ac[j]) * vc[j], f1, -(bc[j] * vc[j]) * f0);
                rr_sts(f_sh + 8u * (unsigned int)(i + j + 1), fn);
                f0 = f1;
                f1 = fn;
            }
        }
    }
}
// returns false if the recurrence overflowed (the caller then uses the rescaling global-memory version)
__device__ __forceinline__ bool rr_three_term_up_smem(unsigned int a_sh, unsigned int b_sh, unsigned int binv_sh, unsigned int g_sh, int k,
                                                      double theta) {
    double g0 = 1.0;   // g_{k-1}
    rr_sts(g_sh + 8u * (unsigned int)(k - 1), 1.0);
    if (k < 2) return true;
    double g1 = -(rr_lds(a_sh + 8u * (unsigned int)(k - 1)) - theta) * rr_lds(binv_sh + 8u * (unsigned int)(k - 1));   // g_{k-2}
    rr_sts(g_sh + 8u * (unsigned int)(k - 2), g1);
    double an[4], bn[4], vn[4];
    auto fetch = [&](int i) {   // rows i, i-1, i-2, i-3 (indices clamped at 0)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const unsigned int r = (unsigned int)max(i - j, 0);
            an[j] = rr_lds(a_sh + 8u * r);
            bn[j] = rr_lds(b_sh + 8u * (r + 1u));
            vn[j] = rr_lds(binv_sh + 8u * r);
        }
    };
    fetch(k - 2);
    for (int i = k - 2; i >= 1; i -= 4) {
        double ac[4], bc[4], vc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { ac[j] = an[j]; bc[j] = bn[j]; vc[j] = vn[j]; }
        fetch(i - 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i - j >= 1) {
                const double gn = fma((theta - ac[j]) * vc[j], g1, -(bc[j] * vc[j]) * g0);   // g_{i-j-1}
                rr_sts(g_sh + 8u * (unsigned int)(i - j - 1), gn);
                g0 = g1;
                g1 = gn;
            }
        }
    }
    return fabs(g1) < 1e250 && fabs(g0) < 1e250;
}

// True iff T_k has an eigenvalue below x: a sign change in the sequence of leading principal minors p_i(x) of T - x I,
//     p_0 = 1,  p_1 = a_0 - x,  p_{i+1} = (a_i - x) p_i - b_i^2 p_{i-1}
// (the pivot of the LDL^T factorisation is q_i = p_{i+1} / p_i; a zero minor counts as a negative pivot).  Division-free: two
// dependent fused multiply-adds per row, ~20 cycles, where the pivot form q_i = (a_i - x) - b_i^2 / q_{i-1} is a chain of
// double-precision divisions (~250 cycles each as an IEEE division, ~100 with a hardware reciprocal seed and two Newton steps):
// with the pivot form a check of T_190 took 115 - 180 us and the solver overshot its stopping point by 40 - 60 % (measured).
// The minors are rescaled every fourth row; the shifts this is used for only have to locate theta to ~1e-7 (the accepted
// pair is polished by a Rayleigh quotient + a second twisted factorisation).  a, b2 are written by this CTA during the same
// launch: plain pointers -- no __restrict__/const, which would let the compiler use the non-coherent read-only path.
__device__ __forceinline__ bool rr_eig_below(double* a, double* b2, int k, double x, double pivmin) {
    (void)pivmin;
    double p0 = 1.0, p1 = a[0] - x;
    if (!(p1 > 0.0)) return true;
    for (int i = 1; i < k; ++i) {
        const double pn = fma(a[i] - x, p1, -b2[i] * p0);
        if (!(pn > 0.0) != !(p1 > 0.0) || pn == 0.0) return true;   // (p1 > 0 is an invariant: we leave at the first change)
        p0 = p1;
        p1 = pn;
        if ((i & 3) == 0) {
            const double m = fabs(p1);
            if (m > 1e100) {
                p0 *= 1e-100;
                p1 *= 1e-100;
            } else if (m < 1e-100) {
                p0 *= 1e100;
                p1 *= 1e100;
            }
        }
    }
    return false;
}

// The same count with T_k in SHARED memory (32-bit shared addresses of a[] and b2[], both 16-byte aligned): four rows per trip,
// their entries fetched with two 16-byte loads per array one trip ahead of their use, so that the chain is the two fused
// multiply-adds per row and nothing else (~20 cycles per row instead of ~70 with generic loads inside the chain, measured as
// 13 000 cycles per round on T_190).  Sign changes are collected from the sign bits of consecutive minors (integer pipe); the
// trip is left at its end, the minors are rescaled there.
__device__ __forceinline__ void rr_lds4(unsigned int addr, double (&v)[4]) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[0]), "=d"(v[1]) : "r"(addr));
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v[2]), "=d"(v[3]) : "r"(addr + 16u));
}
__device__ __forceinline__ bool rr_eig_below_smem(unsigned int a_addr, unsigned int b2_addr, int k, double x) {
    // rows are padded to a multiple of four by the caller: a = +huge, b2 = 0 beyond k keep the sign of the minors
    double av[4], bv[4], an[4], bn[4];
    rr_lds4(a_addr, av);
    rr_lds4(b2_addr, bv);
    double p0 = 1.0, p1 = av[0] - x;
    int flip = __double2hiint(p1);                  // sign bit set <=> p_1 < 0
    if (p1 == 0.0) return true;
    {   // rows 1..3 of the first trip
        rr_lds4(a_addr + 32u, an);
        rr_lds4(b2_addr + 32u, bn);
#pragma unroll
        for (int j = 1; j < 4; ++j) {
            const double pn = fma(av[j] - x, p1, -bv[j] * p0);
            flip |= __double2hiint(pn) ^ __double2hiint(p1);
            p0 = p1;
            p1 = pn;
        }
    }
    if (flip < 0) return true;
    for (int i = 4; i < k; i += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { av[j] = an[j]; bv[j] = bn[j]; }
        rr_lds4(a_addr + 8u * (unsigned int)(i + 4), an);     // (the arrays are padded by four more entries)
        rr_lds4(b2_addr + 8u * (unsigned int)(i + 4), bn);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double pn = fma(av[j] - x, p1, -bv[j] * p0);
            flip |= __double2hiint(pn) ^ __double2hiint(p1);
            p0 = p1;
            p1 = pn;
        }
        if (flip < 0 || !(fabs(p1) < 1e300)) return true;   // (overflow / NaN only after a zero minor: counts as a change)
        const double m = fabs(p1);
        if (m > 1e100) {
            p0 *= 1e-100;
            p1 *= 1e-100;
        } else if (m < 1e-100) {
            p0 *= 1e100;
            p1 *= 1e100;
        }
    }
    return false;
}

__device__ __forceinline__ int rr_next_check(int k, int div) { return k + max(div >= 24 ? 4 : 16, (k / div) & ~3); }

__device__ __noinline__ void lz_rr_main(const RrArgs& R, double* smem, int smem_doubles) {
    __shared__ double red[32];
    // T_k's diagonal and squared off-diagonal for the Sturm counts: in this CTA's (otherwise unused) dynamic shared memory
    // while they fit, in the private global arrays beyond that
    // (every sequential recurrence below is a chain of dependent loads otherwise: an L1 miss to L2 costs 700+ cycles while the
    // solver CTAs are gathering -- 78 us per check with the arrays in global memory, measured)
    const int cap_s = (smem_doubles / 8) & ~3;   // multiple of 4: every array stays 32-byte aligned
    double* const a_s = smem;
    double* const b2_s = smem + cap_s;
    double* const b_s = smem + 2 * cap_s;
    double* const binv_s = smem + 3 * cap_s;
    double* const w_s = smem + 4 * cap_s;   // f (top-down), g (bottom-up), s
    __shared__ int s_fail, s_kbreak, s_kr;
    __shared__ double a_pad_save[8], b2_pad_save[8];
    __shared__ double red4[128];
    const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
    const double lnorm = R.sc->lnorm, sqrtn = sqrt((double)R.n);
    const double brk = fmax(1e-12 * lnorm, 0.25 * R.tol * lnorm / sqrtn);
    const double target = R.tol * lnorm / (0.88 * sqrtn);
    const double eps = 2.220446049250313e-16, dmin = 2.2250738585072014e-308;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    int k_seen = 0, k_next = min(16, R.k_limit), checks = 0, k_prev = 0;
    double theta_prev = inf, theta_delta = -1.0, est_prev = 0.0;
    if (tid == 0) {
        s_fail = 0;
        s_kbreak = 1 << 30;
    }
    const long long t_begin = clock64();
    __syncthreads();
    int status = 0, k = 0;
    long long cyc_wait = 0, cyc_stage[6] = {0, 0, 0, 0, 0, 0};
    double gl = inf, gu = -inf, amin = inf, bmax = 0.0;   // running bounds of T_k
    bool no_freeze = false;
    int k_bounds = 0;
    int rounds = 0;
    double theta = 0.0, est = inf, s_inv = 0.0;
    bool exhausted = false, s_vok = false;
    while (true) {
        const int need = min(k_next, R.k_limit);
        // ---- new coefficients.  ONE thread waits for the newest entry, slowly (a thousand threads polling the two lines solver
        // CTA 0 writes every phase slow that CTA down, and with it the whole grid: 31 us per step instead of 7.5, measured);
        // then every entry is waited for on its own: the solver publishes them with plain stores, nothing orders them
        const long long tw0 = clock64();
        if (tid == 0) {
            unsigned int spin = 0;
            while (true) {
                const double be = ld_relaxed_f64(R.beta + need);
                if (be == be) break;
                __nanosleep(1000);
                if ((++spin & 255u) == 0 && clock64() - t_begin > 6000000000ll) {   // ~3 s: the solver is gone
                    s_fail = 3;
                    break;
                }
            }
        }
        __syncthreads();
        cyc_wait += clock64() - tw0;
        long long ts = clock64();
#define RR_STAGE(i) do { const long long tn_ = clock64(); cyc_stage[i] += tn_ - ts; ts = tn_; } while (0)
        for (int j = k_seen + tid; j <= need; j += nt) {
            double al, be;
            unsigned int spin = 0;
            while (true) {
                al = ld_relaxed_f64(R.alpha + j);
                be = ld_relaxed_f64(R.beta + j);
                if (al == al && be == be) break;
                __nanosleep(200);
                if ((++spin & 1023u) == 0 && clock64() - t_begin > 6000000000ll) {   // ~3 s: the solver is gone
                    atomicExch(&s_fail, 3);
                    al = be = 0.0;
                    break;
                }
                if (*(volatile int*)&s_fail) break;
            }
            if (!(fabs(al) < inf) || !(fabs(be) < inf)) atomicMax(&s_fail, 1);
            R.a[j] = al;
            R.b[j] = be;
            R.b2[j] = be * be;
            if (j < cap_s) {
                a_s[j] = al;
                b2_s[j] = be * be;
                b_s[j] = be;
                binv_s[j] = (be != 0.0) ? 1.0 / be : 0.0;
            }
            R.binv[j] = (be != 0.0) ? 1.0 / be : 0.0;
            if (j >= 1 && !(be > brk)) atomicMin(&s_kbreak, j);   // breakdown: span(u_0..u_{j-1}) is invariant, T_j is exact
        }
        __syncthreads();
        if (s_fail) {
            status = (s_fail == 3) ? -3 : -1;
            break;
        }
        k_seen = max(k_seen, need + 1);
        const bool invariant = s_kbreak <= need;
        k = invariant ? s_kbreak : need;
        if (k == 0) {
            status = -1;
            break;
        }
        ++checks;
        RR_STAGE(0);
        // ---- bounds: Gershgorin, smallest diagonal entry, largest squared off-diagonal -- running values, only the rows that
        // are new (or whose lower neighbour is) are looked at; a row's older, smaller radius stays in the min / max harmlessly
        if (tid < 32) {   // a handful of new rows per check: one warp
            double v0 = inf, v1 = -inf, v2 = inf, v3 = 0.0;
            for (int i = max(k_bounds - 1, 0) + tid; i < k; i += 32) {
                // (shared copies while they reach: a dependent global load costs ~2 500 cycles under the solver's traffic)
                const bool sm_ok = i + 1 < cap_s;
                const double ai = sm_ok ? a_s[i] : ld_relaxed_f64(R.a + i);
                const double bl = (i > 0) ? fabs(sm_ok ? b_s[i] : ld_relaxed_f64(R.b + i)) : 0.0;
                const double bu = (i + 1 < k) ? fabs(sm_ok ? b_s[i + 1] : ld_relaxed_f64(R.b + i + 1)) : 0.0;
                v0 = fmin(v0, ai - (bl + bu));
                v1 = fmax(v1, ai + (bl + bu));
                v2 = fmin(v2, ai);
                if (i > 0) v3 = fmax(v3, bl * bl);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                v0 = fmin(v0, __shfl_xor_sync(0xffffffffu, v0, o));
                v1 = fmax(v1, __shfl_xor_sync(0xffffffffu, v1, o));
                v2 = fmin(v2, __shfl_xor_sync(0xffffffffu, v2, o));
                v3 = fmax(v3, __shfl_xor_sync(0xffffffffu, v3, o));
            }
            if (tid == 0) {
                red4[0] = v0; red4[1] = v1; red4[2] = v2; red4[3] = v3;
            }
        }
        __syncthreads();
        gl = fmin(gl, red4[0]); gu = fmax(gu, red4[1]); amin = fmin(amin, red4[2]); bmax = fmax(bmax, red4[3]);
        k_bounds = k;
        const double pivmin = fmax(dmin, dmin * bmax) * 4.0;
        const double tnorm = fmax(fabs(gl), fabs(gu));
        const bool in_s = k + 9 <= cap_s;
        if (in_s && tid < 8) {   // padding rows for the four-row trips of the Sturm count: they keep the sign of the minors
            a_pad_save[tid] = a_s[k + tid];
            b2_pad_save[tid] = b2_s[k + tid];
        }
        __syncthreads();
        if (in_s && tid < 8) {
            a_s[k + tid] = 1e30;
            b2_s[k + tid] = 0.0;
        }
        __syncthreads();
        const unsigned int a_sh = (unsigned int)__cvta_generic_to_shared(a_s), b2_sh = (unsigned int)__cvta_generic_to_shared(b2_s);
        double* const pa = in_s ? a_s : R.a;
        double* const pb2 = in_s ? b2_s : R.b2;
        double* const pb = in_s ? b_s : R.b;
        double* const pbinv = in_s ? binv_s : R.binv;
        double* const pdp = in_s ? w_s : R.dp;
        double* const pdm = in_s ? w_s + cap_s : R.dm;
        double* const ps = in_s ? w_s + 2 * cap_s : R.s;
        double lo = gl - 2.0 * eps * tnorm * k - 2.0 * pivmin;
        double hi = fmin(gu + 2.0 * eps * tnorm * k + 2.0 * pivmin, amin + 4.0 * eps * tnorm);
        RR_STAGE(1);
        // Once two consecutive checks agree on theta to 2e-7 it is no longer searched for: the eigenvector recurrences below
        // tolerate that error (the tail of the vector moves by (k - i0) d(theta) / gap), and an accepted pair is polished by
        // its Rayleigh quotient.  Ritz values only decrease with k, by less every check.
        // Two attempts.  The first is the cheap one (frozen or 1e-7 theta + Rayleigh-quotient polish); it is checked -- the
        // polish must not move theta by more than 1e-5, a frozen theta must not make the estimate jump up -- and if it fails
        // the check is repeated with a full-precision multisection from the safe bracket, and theta is never frozen again in
        // this solve.  (Without the second attempt one solve in ~60 at the headline size never saw its estimate pass and ran
        // to the end of the basis: 2 700 steps instead of 250, measured.)
        const double lo0 = lo, hi0 = hi;
        double theta0 = 0.0;
        for (int attempt = 0; attempt < 2; ++attempt) {
        const bool precise = attempt == 1;
        lo = lo0;
        hi = hi0;
        const bool theta_frozen = !precise && !no_freeze && theta_prev < inf && theta_delta >= 0.0 &&
                                  theta_delta < 4e-7 * fabs(theta_prev) && !invariant;
        if (k == 1) {
            theta = in_s ? a_s[0] : R.a[0];
        } else if (theta_frozen) {
            theta = theta_prev;
        } else {
            // ---- warm bracket: probes h, h - d, h - 4 d, h - 16 d, ... below the previous theta
            if (theta_prev < inf && !invariant) {
                const double h = theta_prev + 8.0 * eps * tnorm;
                double d = (theta_delta > 0.0) ? theta_delta : 1e-6 * fmax(fabs(h), 1e-300);
                d = fmax(d, 16.0 * eps * tnorm);
                bool neg = false;
                double x = h;
                if (tid < 64) {
                    if (tid > 0) x = h - d * exp2(2.0 * (double)(tid - 1));
                    if (x > lo && x < hi) neg = in_s ? rr_eig_below_smem(a_sh, b2_sh, k, x) : rr_eig_below(pa, pb2, k, x, pivmin);
                    else if (x >= hi) neg = true;    // hi is an upper bound of the eigenvalue
                }
                const int cnt = __syncthreads_count(neg);   // probes 0 .. cnt-1 lie above the eigenvalue (monotone)
                if (cnt > 0) {
                    const double xh = (cnt - 1 == 0) ? h : h - d * exp2(2.0 * (double)(cnt - 2));
                    const double xl = (cnt >= 64) ? lo : h - d * exp2(2.0 * (double)(cnt - 1));
                    hi = fmin(hi, xh);
                    lo = fmax(lo, xl);
                }
            }
            // ---- multisection: kRrProbes interior shifts per round (more would be bound by the SM's 64 FP64 lanes, not by the
            // latency of the recurrence), down to a relative width of 1e-7: the decision needs the residual estimate to ~10 %,
            // and the tail of the eigenvector moves by (k - i0) d(theta) / gap -- 1e-7 in theta would do
            constexpr int kRrProbes = 256;
            RR_STAGE(2);
            for (int round = 0; round < 40; ++round) {
                ++rounds;
                const double width = hi - lo;
                if (!(width > (precise ? 2.0 * eps : 1e-7) * fmax(fabs(lo), fabs(hi)) + 2.0 * pivmin)) break;
                bool neg = false;
                if (tid < kRrProbes) {
                    const double x = lo + width * ((double)(tid + 1) / (double)(kRrProbes + 1));
                    const bool inside = x > lo && x < hi;
                    neg = inside ? (in_s ? rr_eig_below_smem(a_sh, b2_sh, k, x) : rr_eig_below(pa, pb2, k, x, pivmin)) : (x >= hi);
                }
                const int nonneg = kRrProbes - __syncthreads_count(neg);   // shifts 0 .. nonneg-1 lie below (or at) the eigenvalue
                const double nlo = (nonneg == 0) ? lo : lo + width * ((double)nonneg / (double)(kRrProbes + 1));
                const double nhi = (nonneg == kRrProbes) ? hi : lo + width * ((double)(nonneg + 1) / (double)(kRrProbes + 1));
                const bool stuck = !(nhi - nlo < width);
                lo = fmax(lo, nlo);
                hi = fmin(hi, nhi);
                if (stuck) break;
            }
            theta = 0.5 * (lo + hi);
            RR_STAGE(3);
        }
        theta0 = theta;
        for (int polish = 0; polish < 2; ++polish) {
        // ---- eigenvector of T_k for theta: three-term recurrences from both ends (two threads, side by side), joined at the
        // index r of the largest entry of the one from the top.  (A recurrence from the top alone cannot resolve the tail of a
        // converged pair, and est = |beta_k s_{k-1}| needs exactly that tail -- measured: the estimate stalled at 1e-3 of a pair
        // whose true residual had long passed 1e-9.)
        if (in_s) {
            const unsigned int b_sh = (unsigned int)__cvta_generic_to_shared(b_s), binv_sh = (unsigned int)__cvta_generic_to_shared(binv_s);
            if (tid == 0) rr_three_term_down_smem(a_sh, b_sh, binv_sh, (unsigned int)__cvta_generic_to_shared(pdp), k, theta);
            else if (tid == 32 && !rr_three_term_up_smem(a_sh, b_sh, binv_sh, (unsigned int)__cvta_generic_to_shared(pdm), k, theta))
                rr_three_term_up(pa, pb, pbinv, k, theta, pdm);
        } else {
            if (tid == 0) rr_three_term_down(pa, pb, pbinv, k, theta, pdp);
            else if (tid == 32) rr_three_term_up(pa, pb, pbinv, k, theta, pdm);
        }
        __syncthreads();
        // The rest of the pass -- joining the two recurrences at r = argmax |f_i| (ties -> smallest index), normalisation, the
        // estimate, the Rayleigh quotient -- is a few hundred entries: ONE warp with shuffle reductions.  (Spread over the 1 024
        // threads with block-wide reductions it cost 9 400 cycles per pass, more than the two recurrences: measured.)
        if (tid < 32) {
            double best = -1.0;
            int bi = 0;
            for (int i = tid; i < k; i += 32) {
                const double v = fabs(pdp[i]);
                if (v > best) {
                    best = v;
                    bi = i;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) {
                    best = ob;
                    bi = oi;
                }
            }
            const int r_tw = bi;
            const double fr = pdp[r_tw], gr = pdm[r_tw];
            bool vec_ok = fabs(fr) > 0.0 && fabs(fr) < inf && fabs(gr) > 0.0 && fabs(gr) < inf;
            const double gscale = vec_ok ? fr / gr : 0.0;
            double ssq = 0.0;
            for (int i = tid; i < k; i += 32) {
                const double v = (i <= r_tw) ? pdp[i] : gscale * pdm[i];
                ps[i] = v;
                ssq = fma(v, v, ssq);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
            vec_ok = vec_ok && ssq > 0.0 && ssq < inf;
            const double inv = vec_ok ? 1.0 / sqrt(ssq) : 0.0;
            const double s_last = vec_ok ? ((k - 1 <= r_tw) ? pdp[k - 1] : gscale * pdm[k - 1]) : 0.0;
            double theta_new = theta;
            int leave = (!vec_ok || polish == 1 || k == 1) ? 1 : 0;
            if (!leave) {
                // theta is known to ~1e-7 only (multisection stops there, a frozen theta is the previous check's).  The vector just
                // computed is one step of inverse iteration with that shift, so its Rayleigh quotient is accurate to
                // d(theta) (d(theta) / gap)^2 -- Rayleigh-quotient iteration converges cubically -- and the SECOND pass, with the
                // polished theta, gives the tail of the eigenvector that the estimate needs: an error d(theta) in the shift
                // contaminates the vector with other Ritz vectors (last entries ~0.1) at the level d(theta) / gap, far above the
                // 1e-10 the tail of a converged pair has (measured: with theta to 1e-7 and no polish the estimate stalls at 5e-4).
                __syncwarp();
                double num = 0.0;
                for (int i = tid; i < k; i += 32) {
                    double t = pa[i] * ps[i];
                    if (i > 0) t = fma(pb[i], ps[i - 1], t);
                    if (i + 1 < k) t = fma(pb[i + 1], ps[i + 1], t);
                    num = fma(ps[i], t, num);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) num += __shfl_xor_sync(0xffffffffu, num, o);
                const double theta_rq = num * inv * inv;
                if (precise && !(fabs(theta_rq - theta0) <= 1e-9 * fmax(fabs(theta0), 1e-300))) leave = 1;   // keep the bisected pair
                else theta_new = theta_rq;
            }
            if (tid == 0) {
                red[0] = vec_ok ? fabs(pb[k]) * fabs(s_last) * inv : inf;
                red[1] = inv;
                red[2] = theta_new;
                s_kr = (vec_ok ? 1 : 0) | (leave ? 2 : 0);
            }
        }
        __syncthreads();
        RR_STAGE(4);
        est = red[0];
        s_inv = red[1];
        theta = red[2];
        s_vok = (s_kr & 1) != 0;
        exhausted = invariant || need >= R.k_limit;
        const bool leave_polish = (s_kr & 2) != 0;
        __syncthreads();   // red / s_kr are reused by the next pass
        if (leave_polish) break;
        }   // polish
        {   // is the cheap attempt credible?
            const bool moved = !(fabs(theta - theta0) <= 1e-5 * fmax(fabs(theta0), 1e-300));
            const bool jumped = theta_frozen && est_prev > 0.0 && !(est < 4.0 * est_prev);
            if (precise || k == 1 || (s_vok && !moved && !jumped)) break;
            no_freeze = true;
        }
        }   // attempt
        __syncthreads();
        if (in_s && tid < 8) {
            a_s[k + tid] = a_pad_save[tid];
            b2_s[k + tid] = b2_pad_save[tid];
        }
        __syncthreads();
        {
        const bool vec_ok = s_vok;
        const double inv = s_inv;
        if (est < target || exhausted) {
            status = (est < target) ? 1 : (!vec_ok ? -2 : (invariant ? 3 : 2));
            for (int t = tid; t < k; t += nt) R.coef[t] = vec_ok ? ps[t] * inv * R.binv[t] : 0.0;
            break;
        }
        }
        RR_STAGE(5);
        // ---- next check point: a function of the coefficients only, so that a solve is a pure function of its input: from the
        // geometric decay of the estimate, half-way to the step at which it is predicted to cross the target -- the predicted
        // step itself once that is near -- and never closer than a check takes
        int nk = rr_next_check(need, R.check_div);
        if (est_prev > 0.0 && est > target && est < est_prev && k > k_prev) {
            const double slope = (log(est) - log(est_prev)) / (double)(k - k_prev);
            const double pred = (log(target) - log(est)) / slope;
            const double cap = fmax(16.0, 0.25 * (double)k);
            // never closer than a check takes (its cost grows with k: 4-6 Lanczos steps at k = 190 at the headline size): checks
            // spaced closer than that queue up behind each other, and the decision then lags the coefficients by the sum of
            // their durations (measured: 15 steps) instead of one
            const double min_gap = 4.0 * (double)max(1, (k + 48) / 96);
            // close to the crossing (within three gaps): go for the predicted step itself; further away: half-way, the decay
            // rate is not settled yet
            nk = need + (int)fmax(min_gap, (pred <= 3.0 * min_gap) ? ceil(pred) + 1.0 : fmin(0.5 * pred, cap));
        }
        if (theta_prev < inf) theta_delta = 2.0 * fabs(theta_prev - theta);
        theta_prev = theta;
        k_prev = k;
        est_prev = est;
        k_next = nk;
    }
    __syncthreads();
    if (tid == 0) {
        R.out->status = status;
        R.out->k = k;
        R.out->checks = checks;
        R.out->theta = theta;
        R.out->est = est;
        R.out->target = target;
        R.out->cyc_wait = cyc_wait;
        R.out->cyc_compute = (clock64() - t_begin) - cyc_wait;
        int lag = 0;
        while (k + 1 + lag <= R.k_limit && lag < 4096) {
            const double be = ld_relaxed_f64(R.beta + k + 1 + lag);
            if (be != be) break;
            ++lag;
        }
        R.out->lag = lag;
        R.out->rounds = rounds;
        for (int i = 0; i < 6; ++i) R.out->cyc_stage[i] = cyc_stage[i];
        __threadfence();
        atomicExch(R.dev_stop, 1);
    }
}

struct LzPipeShared {
    double pollsum[4 * 8];
    double coef[8];            // k1, k2, k3, k4, alpha, beta of the phase being finished
    double beta_prev, usum_prev, sigma, inv_n;
    int stop, give_up;
};

#define LZ_BAR(id) asm volatile("bar.sync %0, %1;" ::"n"(id), "n"(kPBlock) : "memory")

// Host contract (setup_persist): every CTA has at most (kPWarps - ceil(ncta / 32)) * 32 rows, so that the last ceil(ncta / 32)
// warps own no rows and can act as the polling warps.
template <int VB>
__global__ void __launch_bounds__(kPBlock, 1) k_lanczos_pipe(LzPersistArgs a, LzJdsArgs J, LzPipeArgs P, RrArgs R) {
    if (blockIdx.x == (unsigned int)a.ncta) {   // the extra CTA of the launch: on-device Rayleigh-Ritz / stop decision
        extern __shared__ double rr_smem[];
        lz_rr_main(R, rr_smem, R.smem_doubles);
        return;
    }
    extern __shared__ double prod[];
    __shared__ double sm[4 * kPWarps];
    __shared__ LzPipeShared sh;
    __shared__ double hsum[256];          // partial row sums of the helper warps (the longest slices are split in two)
    __shared__ int stop_in;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = (int)threadIdx.x;
    int* __restrict__ scol = reinterpret_cast<int*>(prod + J.pos_cap);
    int* __restrict__ sjw = scol + J.slot_cap;
    const int ra = J.row_start[blockIdx.x], rb = J.row_start[blockIdx.x + 1];
    const int sa = a.rp[ra], ns = a.rp[rb] - sa;
    const bool has_row = tid < rb - ra;
    const int row = ra + tid;
    const double* __restrict__ jval = J.jval + sa;
    // column cache: a slot whose weight is zero in this launch (a candidate edge outside the support) is marked inactive
    for (int i = tid; i < ns; i += kPBlock) scol[i] = ld_nc(J.jcol + sa + i) | ((ld_nc(jval + i) == 0.0) ? kLzColMask : 0);
    for (int i = tid; i < J.pos_cap; i += kPBlock) prod[i] = 0.0;   // padding and inactive positions are never written again
    if (tid < kLzSliceTab) sjw[tid] = J.jw[(size_t)blockIdx.x * kLzSliceTab + tid];
    int phase = a.st->phase;
    int cur = a.st->cur;
    if (tid == 0) {
        sh.give_up = 0;
        stop_in = 0;
        sh.sigma = P.sc->shift;
        sh.inv_n = 1.0 / (double)a.n;
        sh.beta_prev = a.st->beta_prev;   // 1/beta of the last completed phase
        sh.usum_prev = a.st->usum_prev;
    }
    const unsigned int ncta = (unsigned int)a.ncta;
    const int rows_warps = (rb - ra + 31) >> 5;
    const int poll_warps = ((int)ncta + 31) >> 5;
    const int first_poll_warp = kPWarps - poll_warps;
    // The longest slices are split in two.  The rows are numbered by decreasing length, so warp 0 has the longest (twice the
    // mean on an Erdos-Renyi graph) and the whole CTA waits for its row sums at the end of pass 2.  The warps that neither own
    // rows nor poll ("helpers", nhw of them) take the second half of the entries of slices 0 .. nhw-1.
    const int nhw = max(0, min(min(first_poll_warp - rows_warps, rows_warps), 8));
    // ONE base pointer for all per-row state in L2, rows of `ld` doubles: [0], [1] the two z buffers (= a.sect[0], a.sect[1]),
    // [2], [3] u_j / u_{j-1} alternating like them (the basis is written with evict-first stores: reading it back is a DRAM
    // round trip, measured), [4] the diagonal of L' = L - sigma I in engine order.  (Never index a.sect[] with a run-time value:
    // the whole parameter struct is then copied to LOCAL memory, 128 bytes per thread = 128 KB per CTA through a 48 KB L1, and
    // every use becomes an L2 round trip -- 2 000 cycles per phase, measured.)
#define LZS (a.sect[0])
#define LZLD ((size_t)a.ld)
    if (has_row) __stcg(LZS + 4 * LZLD + row, J.diag[J.perm[row]] - P.sc->shift);
    {
        double su = 0.0, sz = 0.0;
        if (has_row) {
            su = __ldcg(a.basis + (size_t)phase * a.ld + row);
            sz = __ldcg(LZS + cur * LZLD + row);
            __stcg(LZS + (2 + cur) * LZLD + row, su);
            if (phase > 0) {   // z_{phase-1}, u_{phase-1} back where the loop expects them
                __stcg(LZS + (cur ^ 1) * LZLD + row, __ldcg(P.zprev + row));
                __stcg(LZS + (3 - cur) * LZLD + row, __ldcg(a.basis + (size_t)(phase - 1) * a.ld + row));
            }
        }
        if (warp < rows_warps) {   // block sums of (u.z, sum z, u.u, sum u) for the records of the first phase of this launch
            const double r = warp_sum4(su * sz, sz, su * su, su, lane);
            if ((lane & 7) == 0) sm[(lane >> 3) * kPWarps + warp] = r;
        }
    }
    __syncthreads();

    for (int it = 0; it < a.nphases; ++it) {
        const double* __restrict__ Z = LZS + cur * LZLD;   // z_phase, published by every CTA for its own rows
        const int tag = (phase >> 1) & 1, tag_next = ((phase + 1) >> 1) & 1;
#ifdef MACB_PTIMING
        const long long t_start = clock64();
        long long t_p1 = 0, t_rows = 0, t_bar = 0, tb0 = 0, tb1 = 0, tu1 = 0, tu2 = 0, tu3 = 0;
#endif
        // ---- records of this phase: the last warp finishes the block sums and pushes them into every CTA's inbox
        if (warp == kPWarps - 1) {
            const double x0 = (lane < rows_warps) ? sm[lane] : 0.0, x1 = (lane < rows_warps) ? sm[kPWarps + lane] : 0.0,
                         x2 = (lane < rows_warps) ? sm[2 * kPWarps + lane] : 0.0, x3 = (lane < rows_warps) ? sm[3 * kPWarps + lane] : 0.0;
            const double r = warp_sum4(x0, x1, x2, x3, lane);
            double q0 = __shfl_sync(0xffffffffu, r, 0), q1 = __shfl_sync(0xffffffffu, r, 8),
                   q2 = __shfl_sync(0xffffffffu, r, 16), q3 = __shfl_sync(0xffffffffu, r, 24);
            const double inf = __longlong_as_double(0x7ff0000000000000ll);
            q0 = (q0 == q0) ? q0 : inf; q1 = (q1 == q1) ? q1 : inf; q2 = (q2 == q2) ? fabs(q2) : inf; q3 = (q3 == q3) ? q3 : inf;
            if (atomicOr(&sh.give_up, 0)) q0 = inf;   // poison alpha: the Rayleigh-Ritz side sees a non-finite value and reports it
            if (blockIdx.x == 0) {
                int stop_now = 0;
                if (lane == 0) {
                    if (a.stop) {   // host-mapped flag: fetched asynchronously during the previous phase (a PCIe read is slow)
                        asm volatile("cp.async.wait_all;" ::: "memory");
                        stop_now = *(volatile int*)&stop_in;
                    }
                    if (P.dev_stop) {
                        int ds;
                        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(ds) : "l"(P.dev_stop) : "memory");
                        stop_now |= ds;
                    }
                }
                if (__shfl_sync(0xffffffffu, stop_now, 0)) q2 = -q2;
                if (lane == 0 && a.stop)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned int)__cvta_generic_to_shared(&stop_in)), "l"(a.stop) : "memory");
            }
            q0 = lz_tagged(q0, tag);
            double* const box = J.xrec + (size_t)(phase & 1) * ncta * ncta * 4;   // [reader][writer][4]
            for (unsigned int b = lane; b < ncta; b += 32) st_sector(box + ((size_t)b * ncta + blockIdx.x) * 4, q0, q1, q2, q3);
        }
        // ---- pass 1: products of the CTA's slots with the gathered z_phase
        lz_gather_products_tagged<VB>(Z, scol, jval, ns, tid, tag, prod, &sh.give_up);
        // The loads of the tail (the row's state / the polling lane's record) are issued before the barrier that ends pass 1.
        // (Issued one batch of gathers earlier they change nothing -- measured: the tail is bound by the shared-memory / shuffle
        // pipe the row sums keep full, not by these loads.  See DESIGN.md, section 5.)
        double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0, r4 = 0.0;
        const bool polling = warp >= first_poll_warp;
        const unsigned int pb = (unsigned int)(tid - first_poll_warp * 32);
        const double* const mine = J.xrec + (size_t)(phase & 1) * ncta * ncta * 4 + (size_t)blockIdx.x * ncta * 4;
        if (polling) {
            r0 = __longlong_as_double((long long)tag);
            if (pb < ncta)
                asm volatile("ld.relaxed.gpu.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                             : "=d"(r0), "=d"(r1), "=d"(r2), "=d"(r3) : "l"(mine + (size_t)pb * 4) : "memory");
        } else if (has_row) {
            const double* __restrict__ Sr = LZS + row;
            r4 = __ldcg(Sr + 4 * LZLD);
            r0 = __ldcg(Sr + (2 + cur) * LZLD);
            r1 = __ldcg(Sr + cur * LZLD);
            if (phase > 0) {
                r2 = __ldcg(Sr + (3 - cur) * LZLD);
                r3 = __ldcg(Sr + (cur ^ 1) * LZLD);   // the buffer about to be overwritten still holds z_{phase-1}
            }
        }

        // From here to the update the polling warps and the row warps run DIFFERENT code between the same two CTA barriers,
        // so that what a row thread keeps in registers (its state, requested from L2 before the first barrier) is not live
        // across the register-hungry polling / coefficient code: inlined into one path the allocator spills it to local memory.
        // The two paths reach barriers 2 and 3 at different program points.  That is within PTX's rule for bar.sync (all threads
        // of a WARP execute the same barrier instruction; the branch is warp-uniform) but compute-sanitizer's synccheck reports
        // it; the form with one bar.sync site per barrier was built and measured: +110 bytes of spills, 6.7 -> 8.2 us per step.
        // Everything needed from L2 is requested BEFORE the barrier that ends pass 1: an L2 round trip costs ~2 500 cycles
        // while the other SMs are gathering -- as long as the row sums and the update together.
        if (polling) {
            // ---- polling warps: this lane's record of the exchange (pushed a whole SpMV ago by everybody); one record per lane
            // = ONE L2 round trip.  The last lane-0 then runs the (long, double-precision) coefficient chain for the whole CTA
            // while the others sum their rows: nothing of the reduction is left on the critical path.
            double y0 = r0, y1 = r1, y2 = r2, y3 = r3;
            LZ_BAR(2);
#ifdef MACB_PTIMING
            tb0 = clock64();
#endif
            unsigned int spins = 0;
            while (true) {
                const bool ok = (pb >= ncta) || lz_tag_ok(y0, tag);
                if (__all_sync(0xffffffffu, ok)) break;
                if (++spins > (1u << 18)) {
                    atomicExch(&sh.give_up, 1);
                    break;
                }
                if (!ok)
                    asm volatile("ld.relaxed.gpu.global.v4.f64 {%0,%1,%2,%3}, [%4];"
                                 : "=d"(y0), "=d"(y1), "=d"(y2), "=d"(y3) : "l"(mine + (size_t)pb * 4) : "memory");
            }
            if (pb == 0) sh.stop = (__double_as_longlong(y2) < 0) ? 1 : 0;
            if (pb >= ncta) y0 = 0.0;
            y2 = fabs(y2);
            const double t = warp_sum4(y0, y1, y2, y3, lane);
            double P1, P2, P3, P4;
            if (poll_warps == 1) {   // up to 32 CTAs (pose graphs): one polling warp, the totals are already in its lanes
                P1 = __shfl_sync(0xffffffffu, t, 0); P2 = __shfl_sync(0xffffffffu, t, 8);
                P3 = __shfl_sync(0xffffffffu, t, 16); P4 = __shfl_sync(0xffffffffu, t, 24);
            } else {
                if ((lane & 7) == 0) sh.pollsum[(lane >> 3) * 8 + (warp - first_poll_warp)] = t;
                asm volatile("bar.sync 1, %0;" ::"r"(poll_warps * 32) : "memory");
                P1 = P2 = P3 = P4 = 0.0;
                if (tid == kPBlock - 32)
                    for (int w = 0; w < poll_warps; ++w) {   // fixed order: identical totals on every CTA
                        P1 += sh.pollsum[w]; P2 += sh.pollsum[8 + w]; P3 += sh.pollsum[16 + w]; P4 += sh.pollsum[24 + w];
                    }
            }
            if (tid == kPBlock - 32) {
                const LzCoef c0 = lz_coefficients(P1, P2, P3, P4, (phase > 0) ? sh.beta_prev : 0.0, sh.usum_prev, sh.inv_n);
                sh.coef[0] = c0.k1; sh.coef[1] = c0.k2; sh.coef[2] = c0.k3; sh.coef[3] = c0.k4;
                sh.coef[4] = c0.alpha; sh.coef[5] = c0.beta;
                sh.beta_prev = c0.binv;
                sh.usum_prev = P4;
            }
#ifdef MACB_PTIMING
            tb1 = clock64();
#endif
            LZ_BAR(3);
        } else {
            // ---- row warps (and helpers): the row's state from L2 (u_j, u_{j-1}, z_j, z_{j-1}, diagonal; not carried in
            // registers across the gather loop, where it would be spilled), the row sums, the update
            const double su = r0, sz = r1, sq = r2, szp = r3, od = r4;
            double qr = 0.0;
            LZ_BAR(2);
#ifdef MACB_PTIMING
            t_p1 = clock64();
#endif
            // ---- pass 2: the products of the CTA's rows, summed along their slices
            bool helped = false;
            if (warp < rows_warps) {
                const int2 wl = *reinterpret_cast<const int2*>(sjw + 2 * warp);   // first position, entries per row
                helped = warp < nhw;
                qr = lz_slice_sum(prod + wl.x + lane, 0, helped ? ((wl.y + 1) >> 1) : wl.y);
            } else if (warp - rows_warps < nhw) {
                const int hw = warp - rows_warps;
                const int2 wl = *reinterpret_cast<const int2*>(sjw + 2 * hw);
                hsum[hw * 32 + lane] = lz_slice_sum(prod + wl.x + lane, (wl.y + 1) >> 1, wl.y);
            }
#ifdef MACB_PTIMING
            t_rows = clock64();
#endif
            LZ_BAR(3);
#ifdef MACB_PTIMING
            t_bar = clock64();
#endif
            // ---- update of the CTA's rows, block sums for the next records
            double p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0;
            if (has_row) {
                const double k1 = sh.coef[0], k2 = sh.coef[1], k3 = sh.coef[2], k4 = sh.coef[3];
                if (helped) qr += hsum[tid];
                qr = fma(od, sz, -qr);   // (L' z_phase)[row]
                const double un = fma(k1, sz, fma(k2, su, k3 * sq)) + k4;
                const double zn = lz_tagged(fma(k1, qr, fma(k2, sz, k3 * szp)) - sh.sigma * k4, tag_next);
#ifdef MACB_PTIMING
                asm volatile("mov.u64 %0, %%clock64;" : "=l"(tu1) : "d"(zn), "d"(un));
#endif
                __stcg(LZS + (cur ^ 1) * LZLD + row, zn);
                __stcg(LZS + (3 - cur) * LZLD + row, un);
                __stcs(a.basis + (size_t)(phase + 1) * a.ld + row, un);   // streaming: the basis must not push the matrix out of L2
                p1 = un * zn; p2 = zn; p3 = un * un; p4 = un;
#ifdef MACB_PTIMING
                tu2 = clock64();
#endif
            }
            if (warp < rows_warps) {
                const double r = warp_sum4(p1, p2, p3, p4, lane);
                if ((lane & 7) == 0) sm[(lane >> 3) * kPWarps + warp] = r;
#ifdef MACB_PTIMING
                asm volatile("mov.u64 %0, %%clock64;" : "=l"(tu3) : "d"(r));
#endif
            }
            if (blockIdx.x == 0 && tid == 0) {
                const double alpha = sh.coef[4] + sh.sigma, beta = sh.coef[5];
                a.alpha[phase] = alpha;
                a.beta[phase] = beta;
                if (a.ab_host)
                    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(a.ab_host + 2 * (size_t)phase), "d"(alpha), "d"(beta) : "memory");
            }
        }
#ifdef MACB_PTIMING
        if ((tid == 0 || tid == kPBlock - 32) && a.timing && it < 64) {
            long long* t = a.timing + ((size_t)it * a.ncta + blockIdx.x) * 8;
            if (tid == 0) { t[0] = t_start; t[1] = t_p1; t[2] = t_rows; t[6] = t_bar; t[7] = clock64(); t[3] = tu1; t[4] = tu2; t[5] = tu3; }
            (void)tb0; (void)tb1;
        }
#endif
        cur ^= 1;
        ++phase;
        __syncthreads();   // sm[] complete for the last warp; sh.coef / sh.stop free for the next phase
        if (sh.stop) break;
    }
    // z_{phase-1} sits in the buffer the next update would overwrite; a later launch (resume) finds it in zprev
    if (has_row && phase > 0) P.zprev[row] = __ldcg(LZS + (cur ^ 1) * LZLD + row);
    if (blockIdx.x == 0 && tid == 0) {
        a.st->phase = phase;
        if (a.stop) const_cast<int*>(a.stop)[1] = phase;   // host-mapped: read after the stream synchronise, no extra copy
        a.st->cur = cur;
        a.st->beta_prev = sh.beta_prev;
        a.st->usum_prev = sh.usum_prev;
        if (R.enabled) R.out->phases = phase;
    }
#undef LZS
#undef LZLD
}

// ---- K3, single-CTA form for small graphs (pose graphs with n up to 3072 nodes, 12288 slots) -------------
// When 24 n + 16 nnz bytes fit in one SM's shared memory the whole problem lives on that SM: (z, u, u') per
// node, the edge weights and the per-slot products in shared memory; column indices, row extents and the
// diagonal in registers (each thread owns slots tid + 1024 j and rows tid + 1024 r for the whole kernel).  A step
// is two block-level passes and __syncthreads -- no grid barrier and no global load at all on the critical
// path.  Same recurrence and outputs (basis, alpha, beta, streamed alpha/beta, stop flag) as k_lanczos_slots.
constexpr int kSmallSlots = 12;   // slots per thread  => nnz <= 12 * 1024
constexpr int kSmallRows = 3;     // rows per thread   => n   <=  3 * 1024

// ---- K3, single-CTA form, second generation (default for small graphs) ------------------------------------------
// Organised like the grid kernel: the current Lanczos vector is materialised in
// shared memory, every thread keeps (u_j, u_{j-1}) of its rows in registers, the block-wide sums are finished by EVERY
// warp redundantly (one shared-memory round instead of a second stage + barrier), the coefficient chain runs on all
// threads.  Three CTA barriers per step instead of six.  State hand-over through the sector buffer is compatible with
// k_lz_persist_init and with itself: (0, u_j, u_{j-1}, diag) with coefficients (0, 1, 0, 0).
__global__ void __launch_bounds__(kPBlock, 1) k_lanczos_small2(LzPersistArgs a, const double* __restrict__ diag, RrArgs R) {
    extern __shared__ double smem_small[];
    if (blockIdx.x == 1) {   // second CTA of the launch: on-device Rayleigh-Ritz / stop decision (see lz_rr_main)
        lz_rr_main(R, smem_small, R.smem_doubles);
        return;
    }
    const int n = a.n;
    const int nnz = a.rp[n];
    double* __restrict__ uvec = smem_small;            // [n]   u_phase
    double* __restrict__ prod = uvec + n;              // [nnz]
    double* __restrict__ wts = prod + nnz;             // [nnz]
    __shared__ double sm[4 * kPWarps];
    __shared__ int stop_small;
    const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) stop_small = 0;

    int pc[kSmallSlots];
#pragma unroll
    for (int j = 0; j < kSmallSlots; ++j) {
        const int i = tid + kPBlock * j;
        pc[j] = (i < nnz) ? a.col[i] : 0;
        if (i < nnz) wts[i] = a.val[i];
    }
    int rs0[kSmallRows], rs1[kSmallRows];
    double rd[kSmallRows], su[kSmallRows], sq[kSmallRows];
    int phase = a.st->phase;
    double* __restrict__ G = a.sect[0];
    {
        const double k1 = a.st->k1, k2 = a.st->k2, k3 = a.st->k3, k4 = a.st->k4;
#pragma unroll
        for (int r = 0; r < kSmallRows; ++r) {
            const int row = tid + kPBlock * r;
            rs0[r] = (row < n) ? a.rp[row] : 0;
            rs1[r] = (row < n) ? a.rp[row + 1] : 0;
            rd[r] = (row < n) ? diag[row] : 0.0;
            su[r] = 0.0;
            sq[r] = 0.0;
            if (row < n) {
                double z, u, q, d;
                ld_sector(G + 4 * (size_t)row, z, u, q, d);
                su[r] = fma(k1, z, fma(k2, u, k3 * q)) + k4;   // u_phase (a sector engine may have left z != 0)
                sq[r] = u;
                if (k1 == 0.0 && k2 == 1.0 && k3 == 0.0) sq[r] = q;   // own hand-over format: (0, u_j, u_{j-1}, diag)
                uvec[row] = su[r];
                if (phase == 0) a.basis[row] = su[r];
            }
        }
    }
    double beta_prev = a.st->beta_prev, usum_prev = a.st->usum_prev;   // beta_prev: 1/beta of the last completed step
    const double inv_n = 1.0 / (double)a.n;
    int stop_probe = 0;
    bool stop_all = false;
    __syncthreads();

    for (int it = 0; it < a.nphases && !stop_all; ++it) {
        // host stop flag: requested every 8th phase, looked at 7 phases later (a host-memory load takes microseconds)
        if (tid == 0 && a.stop && (it & 7) == 0)
            asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(stop_probe) : "l"(a.stop));
        // device-resident flag of the Rayleigh-Ritz CTA: requested every 4th phase, looked at 3 phases later
        if (tid == 0 && R.enabled && (it & 3) == 0)
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(stop_probe) : "l"(R.dev_stop) : "memory");
        // ---- pass 1: products from the shared-memory vector
#pragma unroll
        for (int j = 0; j < kSmallSlots; ++j) {
            const int i = tid + kPBlock * j;
            if (i < nnz) {
                const double w = wts[i];
                prod[i] = (w != 0.0) ? w * uvec[pc[j]] : 0.0;
            }
        }
        __syncthreads();
        // ---- pass 2: z for the thread's rows, partial sums
        double zr[kSmallRows];
        double p1 = 0.0, p2 = 0.0, p3 = 0.0, p4 = 0.0;
#pragma unroll
        for (int r = 0; r < kSmallRows; ++r) {
            const int row = tid + kPBlock * r;
            zr[r] = 0.0;
            if (row < n) {
                double acc0 = 0.0, acc1 = 0.0;
                int i = rs0[r];
                for (; i + 1 < rs1[r]; i += 2) {
                    acc0 += prod[i];
                    acc1 += prod[i + 1];
                }
                if (i < rs1[r]) acc0 += prod[i];
                zr[r] = fma(rd[r], su[r], -(acc0 + acc1));
                p1 = fma(su[r], zr[r], p1);
                p2 += zr[r];
                p3 = fma(su[r], su[r], p3);
                p4 += su[r];
            }
        }
        {
            const double rsum = warp_sum4(p1, p2, p3, p4, lane);
            if ((lane & 7) == 0) sm[(lane >> 3) * kPWarps + warp] = rsum;
        }
        if (tid == 0 && a.stop && (it & 7) == 7) stop_small = stop_probe;
        if (tid == 0 && R.enabled && (it & 3) == 3) stop_small = stop_probe;
        __syncthreads();
        // ---- every warp finishes the block sums itself, then the coefficient chain on every thread
        const double t4 = warp_sum4(sm[lane], sm[kPWarps + lane], sm[2 * kPWarps + lane], sm[3 * kPWarps + lane], lane);
        const double P1 = __shfl_sync(0xffffffffu, t4, 0), P2 = __shfl_sync(0xffffffffu, t4, 8),
                     P3 = __shfl_sync(0xffffffffu, t4, 16), P4 = __shfl_sync(0xffffffffu, t4, 24);
        const LzCoef cf = lz_coefficients(P1, P2, P3, P4, (phase > 0) ? beta_prev : 0.0, usum_prev, inv_n);
        stop_all = (stop_small != 0);
        // ---- pass B: u_{phase+1} for the thread's rows
        double* __restrict__ bn = a.basis + (size_t)(phase + 1) * a.ld;
#pragma unroll
        for (int r = 0; r < kSmallRows; ++r) {
            const int row = tid + kPBlock * r;
            if (row < n) {
                const double un = fma(cf.k1, zr[r], fma(cf.k2, su[r], cf.k3 * sq[r])) + cf.k4;
                uvec[row] = un;
                __stcs(bn + row, un);
                sq[r] = su[r];
                su[r] = un;
            }
        }
        if (tid == 0) {
            a.alpha[phase] = cf.alpha;
            a.beta[phase] = cf.beta;
            if (a.ab_host)
                asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(a.ab_host + 2 * (size_t)phase), "d"(cf.alpha), "d"(cf.beta) : "memory");
        }
        beta_prev = cf.binv;
        usum_prev = P4;
        ++phase;
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < kSmallRows; ++r) {
        const int row = tid + kPBlock * r;
        if (row < n) st_sector(G + 4 * (size_t)row, 0.0, su[r], sq[r], rd[r]);
    }
    if (tid == 0) {
        a.st->phase = phase;
        if (a.stop) const_cast<int*>(a.stop)[1] = phase;   // host-mapped: read after the stream synchronise, no extra copy
        a.st->cur = 0;
        a.st->k1 = 0.0; a.st->k2 = 1.0; a.st->k3 = 0.0; a.st->k4 = 0.0;
        a.st->beta_prev = beta_prev;
        a.st->usum_prev = usum_prev;
        if (R.enabled) R.out->phases = phase;
    }
}


// sectors for phase 0: (0, src_i, 0, diag_i) with (k1,k2,k3,k4) = (0,1,0,0)  =>  u_0 = src
__global__ void __launch_bounds__(kBlock) k_lz_persist_init(int n, const double* __restrict__ src,
                                                            const double* __restrict__ diag, double* __restrict__ sect0,
                                                            LzPersistState* st, LzPartRec* recs, int nrecs,
                                                            const int* __restrict__ perm /* engine -> caller numbering, or null */,
                                                            double* __restrict__ xrec, int64_t nxrec /* inboxes <- NaN */) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nxrec; i += (int64_t)gridDim.x * blockDim.x)
        xrec[i] = __longlong_as_double(0x7ff8000000000000ll);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int o = perm ? perm[i] : i;
        st_sector(sect0 + 4 * (size_t)i, 0.0, src[o], 0.0, diag[o]);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrecs; i += gridDim.x * blockDim.x) recs[i].tag = 0ull;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->phase = 0;
        st->cur = 0;
        st->k1 = 0.0; st->k2 = 1.0; st->k3 = 0.0; st->k4 = 0.0;
        st->beta_prev = 0.0;
        st->usum_prev = 0.0;
        st->bar = 0u;
    }
}

// ---- finalisation: Ritz vector, normalisation, residual ------------------------------------------
// y_raw = sum_t coef[t] basis[t]   (coef[t] = s_t / beta[t] folds the normalisation of u_t)
__global__ void __launch_bounds__(kBlock) k_ritz(int n, int ld, int k, const double* __restrict__ basis,
                                                 const double* __restrict__ coef, double* __restrict__ out,
                                                 LzScalars* sc, ReduceWS ws,
                                                 const int* __restrict__ perm /* engine -> caller numbering, or null */,
                                                 const RrOut* __restrict__ rr /* k decided on the device, or null */) {
    __shared__ double sm[2 * kWarpsPerBlock];
    __shared__ double part[kWarpsPerBlock][128 + 4];
    __shared__ int flag;
    if (rr) k = rr->k;
    // A tile of 128 nodes per block, four consecutive nodes per lane (32-byte loads: every warp reads 1 KB runs of a basis
    // vector -- 256-byte runs left HBM at 2.8 TB/s, ncu), the basis vectors dealt out to the block's warps (warp g takes
    // t = k-1-g, k-1-g-W, ...).  ld is a multiple of 32, so the loads are aligned and stay inside the row.
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    constexpr int W = kWarpsPerBlock;
    double r0 = 0.0, r1 = 0.0;
    const int ntiles = (n + 127) >> 7;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i4 = tile * 128 + 4 * lane;
        double acc[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
        if (i4 < ld) {
            // newest vectors first: the last ~100 MB the Lanczos kernel wrote are still in L2
            int t = k - 1 - grp;
            for (; t >= W; t -= 2 * W) {
                const double2* p0 = reinterpret_cast<const double2*>(basis + (size_t)t * ld + i4);
                const double2* p1 = reinterpret_cast<const double2*>(basis + (size_t)(t - W) * ld + i4);
                const double2 a0 = __ldg(p0), a1 = __ldg(p0 + 1), b0 = __ldg(p1), b1 = __ldg(p1 + 1);
                const double c0 = coef[t], c1 = coef[t - W];
                acc[0][0] = fma(c0, a0.x, acc[0][0]); acc[0][1] = fma(c0, a0.y, acc[0][1]);
                acc[0][2] = fma(c0, a1.x, acc[0][2]); acc[0][3] = fma(c0, a1.y, acc[0][3]);
                acc[1][0] = fma(c1, b0.x, acc[1][0]); acc[1][1] = fma(c1, b0.y, acc[1][1]);
                acc[1][2] = fma(c1, b1.x, acc[1][2]); acc[1][3] = fma(c1, b1.y, acc[1][3]);
            }
            if (t >= 0) {
                const double2* p0 = reinterpret_cast<const double2*>(basis + (size_t)t * ld + i4);
                const double2 a0 = __ldg(p0), a1 = __ldg(p0 + 1);
                const double c0 = coef[t];
                acc[0][0] = fma(c0, a0.x, acc[0][0]); acc[0][1] = fma(c0, a0.y, acc[0][1]);
                acc[0][2] = fma(c0, a1.x, acc[0][2]); acc[0][3] = fma(c0, a1.y, acc[0][3]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) part[grp][4 * lane + q] = acc[0][q] + acc[1][q];
        __syncthreads();
        const int i = tile * 128 + (int)threadIdx.x;
        if (threadIdx.x < 128 && i < n) {
            double yv = 0.0;
#pragma unroll
            for (int g = 0; g < W; ++g) yv += part[g][threadIdx.x];
            out[perm ? perm[i] : i] = yv;
            r0 += yv;
            r1 = fma(yv, yv, r1);
        }
        __syncthreads();
    }
    double v[2] = {r0, r1};
    if (grid_reduce<2>(v, ws, sm, &flag)) {
        sc->ritz_sum = v[0];
        sc->ritz_sq = v[1];
    }
}

// v = (y_raw - mean) / ||y_raw - mean||
__global__ void __launch_bounds__(kBlock) k_center_normalize(int n, double* __restrict__ v, const LzScalars* sc) {
    const double mean = sc->ritz_sum / (double)n;
    const double nrm2 = sc->ritz_sq - (double)n * mean * mean;
    const double inv = (nrm2 > 0.0) ? 1.0 / sqrt(nrm2) : 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = (v[i] - mean) * inv;
}

// res1 = sum_i | (L v)_i - theta v_i |, theta = vLv / vv   (nx:243)
__global__ void __launch_bounds__(kBlock) k_resid_l1(int n, const double* __restrict__ v, const double* __restrict__ lv,
                                                     LzScalars* sc, ReduceWS ws) {
    __shared__ double sm[kWarpsPerBlock];
    __shared__ int flag;
    const double theta = sc->vLv / sc->vv;
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        acc += fabs(lv[i] - theta * v[i]);
    double r[1] = {acc};
    if (grid_reduce<1>(r, ws, sm, &flag)) sc->res1 = r[0];
}

// ---- K4: gradient g_k = kappa_k (v_i - v_j)^2  (mac.py:117-124), fused ||g||^2 and g.x ------------
__global__ void __launch_bounds__(kBlock) k_gradient(int64_t m, const int* __restrict__ ci, const int* __restrict__ cj,
                                                     const double* __restrict__ kappa, const double* __restrict__ v,
                                                     const double* __restrict__ x, double* __restrict__ g,
                                                     LzScalars* sc, ReduceWS ws) {
    __shared__ double sm[2 * kWarpsPerBlock];
    __shared__ int flag;
    double r0 = 0.0, r1 = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (int64_t)gridDim.x * blockDim.x) {
        const double d = ld_nc(v + ld_stream(ci + e)) - ld_nc(v + ld_stream(cj + e));
        const double kd = __dmul_rn(ld_stream(kappa + e), d);   // kdelta = weight_k * (v_i - v_j)
        const double ge = __dmul_rn(kd, d);                     // gradf[k] = kdelta * (v_i - v_j)
        g[e] = ge;
        r0 = fma(ge, ge, r0);
        r1 = fma(ge, x[e], r1);
    }
    double r[2] = {r0, r1};
    if (grid_reduce<2>(r, ws, sm, &flag)) {
        sc->gnorm2 = r[0];
        sc->gdotx = r[1];
    }
}

// ---- K5: top-k by MSD radix select on the order-preserving 64-bit key of g -------------------------
struct SelState {
    unsigned long long prefix;   // key bits decided so far
    unsigned long long mask;     // which bits of `prefix` are decided
    long long remaining;         // how many of the current bucket still have to be taken
    long long count_gt;          // elements strictly above the current bucket
    long long eq_total;          // after the last pass: elements equal to the k-th key
    unsigned int hist[256];
};

__device__ __forceinline__ unsigned long long order_key(double g) {
    unsigned long long b = (unsigned long long)__double_as_longlong(g);
    if (b == 0x8000000000000000ull) b = 0ull;   // -0.0 and +0.0 compare equal (numpy / IEEE), so they share a key
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(kBlock) k_sel_init(SelState* st, long long k) {
    if (threadIdx.x == 0) {
        st->prefix = 0ull;
        st->mask = 0ull;
        st->remaining = k;
        st->count_gt = 0;
        st->eq_total = 0;
    }
    st->hist[threadIdx.x] = 0u;
}

// ---- K5, two-pass form: one 15-bit histogram of the whole array + an exact select inside the chosen bin ------------------------
// The eight (histogram, pick) launch pairs above cost more than the eigen-solve on small graphs (17 launches per LP step) and
// read g eight times.  Here:
//   k_sel2_hist     ONE pass over g: histogram of the top 15 key bits (sign, exponent, three mantissa bits) in shared memory
//                   (32 768 bins = 128 KB, one CTA per SM), non-empty bins merged into a global histogram; the last CTA to finish
//                   walks it from the top and fixes the bin holding the k-th largest key, the rank inside it and the count above;
//   k_sel2_compact  second pass over g: the keys of that bin (a few per cent of the array) appended to a candidate buffer;
//   k_sel2_refine   one CTA: the remaining 49 bits by four 13/13/13/10-bit histogram passes over the candidates.
// Arrays of up to kSel2SmallMax elements do all of it in one single-CTA launch (k_sel2_small).  The result is the same SelState
// (exact 64-bit key of the k-th element, ties to take, count above, number of ties) the apply kernels below consume.
constexpr int kSel2Bins = 1 << 15;
constexpr int kSel2Block = 1024;
constexpr int kSel2SmallMax = 1 << 16;

struct Sel2State {
    unsigned int bin;        // top 15 key bits of the k-th largest key
    unsigned int ncand;      // candidates appended so far (k_sel2_compact)
    long long remaining;     // rank of the k-th largest key inside the bin, counted from the bin's largest (1-based)
    long long count_gt;      // keys in bins above
    unsigned int done;       // CTAs of k_sel2_hist that have merged their histogram
    unsigned int pad;
};

// Block-wide (kSel2Block threads): the bin d, counted from the TOP, in which the cumulative count reaches `rem`
// (above < rem <= above + hist[d]); nbins a multiple of kSel2Block.  Results through shared memory.
__device__ __forceinline__ void sel2_find_from_top(const unsigned int* hist, int nbins, long long rem, unsigned long long* warp_tot /*[32]*/,
                                                   int* out_digit, long long* out_above, unsigned int* out_count) {
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = nbins / kSel2Block;
    unsigned long long tsum = 0;
    for (int b = 0; b < per; ++b) tsum += hist[tid * per + b];
    // inclusive suffix sum inside the warp, then the totals of the warps above
    unsigned long long incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_down_sync(0xffffffffu, incl, o);
        if (lane + o < 32) incl += t;
    }
    if (lane == 0) warp_tot[warp] = incl;
    __syncthreads();
    unsigned long long above = incl - tsum;
    for (int w = warp + 1; w < kSel2Block / 32; ++w) above += warp_tot[w];
    if ((long long)above < rem && (long long)(above + tsum) >= rem) {
        unsigned long long a = above;
        for (int b = per - 1; b >= 0; --b) {
            const unsigned int h = hist[tid * per + b];
            if ((long long)(a + h) >= rem) {
                *out_digit = tid * per + b;
                *out_above = (long long)a;
                *out_count = h;
                break;
            }
            a += h;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void sel2_hist_add(unsigned int* sh, unsigned int d, bool ok) {
    // warp-aggregated update (the leading key bits are nearly constant across the array)
    const unsigned int peers = __match_any_sync(0xffffffffu, ok ? d : 0xffffffffu);
    if (ok && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&sh[d], (unsigned int)__popc(peers));
}

__global__ void __launch_bounds__(kSel2Block, 1) k_sel2_hist(int64_t m, const double* __restrict__ g, long long k, unsigned int* __restrict__ ghist,
                                                              Sel2State* s2) {
    extern __shared__ unsigned int sh2[];   // [kSel2Bins]
    __shared__ unsigned long long warp_tot[32];
    __shared__ int digit;
    __shared__ long long above_s;
    __shared__ unsigned int count_s;
    __shared__ int last;
    for (int i = threadIdx.x; i < kSel2Bins; i += kSel2Block) sh2[i] = 0u;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * kSel2Block;
    const int64_t mceil = ((m + 31) / 32) * 32;
    // eight loads in flight per thread: one load per trip of a loop with a warp-synchronous step in it is a chain of L2 latencies
    for (int64_t e0 = (int64_t)blockIdx.x * kSel2Block + threadIdx.x; e0 < mceil; e0 += 8 * stride) {
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (e0 + q * stride < m) ? ld_stream(g + e0 + q * stride) : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (e0 + q * stride >= mceil) break;   // warp-uniform
            const bool ok = e0 + q * stride < m;
            sel2_hist_add(sh2, ok ? (unsigned int)(order_key(v[q]) >> 49) : 0u, ok);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSel2Bins; i += kSel2Block) {
        const unsigned int c = sh2[i];
        if (c) atomicAdd(&ghist[i], c);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(&s2->done, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!last) return;
    __threadfence();
    {   // all 32 loads of the thread first, then the stores: interleaved they serialise on the L2 latency (17 of 27 us, ncu)
        unsigned int hv[kSel2Bins / kSel2Block];
#pragma unroll
        for (int q = 0; q < kSel2Bins / kSel2Block; ++q) hv[q] = __ldcg(ghist + threadIdx.x + q * kSel2Block);
#pragma unroll
        for (int q = 0; q < kSel2Bins / kSel2Block; ++q) {
            sh2[threadIdx.x + q * kSel2Block] = hv[q];
            if (hv[q]) ghist[threadIdx.x + q * kSel2Block] = 0u;   // ready for the next selection
        }
    }
    __syncthreads();
    sel2_find_from_top(sh2, kSel2Bins, k, warp_tot, &digit, &above_s, &count_s);
    if (threadIdx.x == 0) {
        s2->bin = (unsigned int)digit;
        s2->remaining = k - above_s;
        s2->count_gt = above_s;
        s2->ncand = 0u;
        s2->done = 0u;
    }
}

__global__ void __launch_bounds__(kBlock) k_sel2_compact(int64_t m, const double* __restrict__ g, Sel2State* s2,
                                                         unsigned long long* __restrict__ cand) {
    const unsigned int bin = s2->bin;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t mceil = ((m + 31) / 32) * 32;
    const int lane = threadIdx.x & 31;
    for (int64_t e0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e0 < mceil; e0 += 4 * stride) {
        double v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = (e0 + q * stride < m) ? ld_stream(g + e0 + q * stride) : 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (e0 + q * stride >= mceil) break;   // warp-uniform
            const unsigned long long key = order_key(v[q]);
            const bool hit = (e0 + q * stride < m) && (unsigned int)(key >> 49) == bin;
            const unsigned int bal = __ballot_sync(0xffffffffu, hit);
            if (bal) {
                unsigned int base = 0;
                if (lane == __ffs(bal) - 1) base = atomicAdd(&s2->ncand, (unsigned int)__popc(bal));
                base = __shfl_sync(0xffffffffu, base, __ffs(bal) - 1);
                if (hit) cand[base + __popc(bal & ((1u << lane) - 1u))] = key;
            }
        }
    }
}

// The low 49 key bits of the k-th largest key: four histogram passes (13, 13, 13, 10 bits) over the candidates.
// FROM_G: the candidates are the elements of g whose top 15 key bits equal the chosen bin (single-CTA path, no compaction).
template <bool FROM_G>
__device__ __forceinline__ void sel2_refine_body(int64_t n, const unsigned long long* __restrict__ cand, const double* __restrict__ g,
                                                 unsigned int bin, long long rem, long long cgt, unsigned int* hist /*[8192]*/,
                                                 unsigned long long* warp_tot, int* digit, long long* above_s, unsigned int* count_s, SelState* st) {
    unsigned long long prefix = (unsigned long long)bin << 49, decided = 0x7fffull << 49;
    unsigned int eq = 0;
    const int shifts[4] = {36, 23, 10, 0};
    const int bits[4] = {13, 13, 13, 10};
    for (int p = 0; p < 4; ++p) {
        const int nb = 1 << bits[p];
        for (int i = threadIdx.x; i < 8192; i += kSel2Block) hist[i] = 0u;
        __syncthreads();
        const int64_t nceil = ((n + 31) / 32) * 32;
        for (int64_t i0 = threadIdx.x; i0 < nceil; i0 += 8 * kSel2Block) {
            unsigned long long kq[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int64_t i = i0 + q * kSel2Block;
                kq[q] = (i < n) ? (FROM_G ? order_key(g[i]) : __ldcg(cand + i)) : 0ull;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int64_t i = i0 + q * kSel2Block;
                if (i >= nceil) break;   // warp-uniform
                const bool ok = (i < n) && (kq[q] & decided) == prefix;
                // plain shared-memory atomics: mantissa digits are all different inside a warp, where the match-any aggregation of
                // sel2_hist_add degenerates into 32 rounds per call (it made this kernel 34 us for 50 000 candidates)
                if (ok) atomicAdd(&hist[(unsigned int)((kq[q] >> shifts[p]) & (unsigned long long)(nb - 1))], 1u);
            }
        }
        __syncthreads();
        sel2_find_from_top(hist, 8192, rem, warp_tot, digit, above_s, count_s);   // bins >= nb are empty
        prefix |= (unsigned long long)(*digit) << shifts[p];
        decided |= (unsigned long long)(nb - 1) << shifts[p];
        cgt += *above_s;
        rem -= *above_s;
        eq = *count_s;
        __syncthreads();
        if (p < 3 && eq <= (unsigned int)kSel2Block) {
            // A handful of candidates left (typically after the first pass): rank them directly instead of three more histogram
            // passes whose cost is all fixed overhead.  One key per thread; every thread counts the keys above and equal to its own.
            unsigned long long* list = reinterpret_cast<unsigned long long*>(hist);   // [kSel2Block]
            unsigned int* cnt = hist + 2 * kSel2Block;
            if (threadIdx.x == 0) *cnt = 0u;
            __syncthreads();
            for (int64_t i0 = threadIdx.x; i0 < n; i0 += 8 * kSel2Block) {
                unsigned long long kq[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int64_t i = i0 + q * kSel2Block;
                    kq[q] = (i < n) ? (FROM_G ? order_key(g[i]) : __ldcg(cand + i)) : 0ull;
                }
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (i0 + q * kSel2Block < n && (kq[q] & decided) == prefix) list[atomicAdd(cnt, 1u)] = kq[q];
            }
            __syncthreads();
            const unsigned int c = *cnt;   // == eq
            if (threadIdx.x < c) {
                const unsigned long long mine = list[threadIdx.x];
                unsigned int gt = 0, same = 0;
                for (unsigned int j = 0; j < c; ++j) {
                    const unsigned long long o = list[j];
                    gt += (o > mine) ? 1u : 0u;
                    same += (o == mine) ? 1u : 0u;
                }
                // exactly the threads holding the k-th largest key satisfy this; they all write the same values
                if ((long long)gt < rem && rem <= (long long)(gt + same)) {
                    st->prefix = mine;
                    st->mask = ~0ull;
                    st->remaining = rem - (long long)gt;
                    st->count_gt = cgt + (long long)gt;
                    st->eq_total = (long long)same;
                }
            }
            return;
        }
    }
    if (threadIdx.x == 0) {
        st->prefix = prefix;
        st->mask = ~0ull;
        st->remaining = rem;
        st->count_gt = cgt;
        st->eq_total = (long long)eq;
    }
}

__global__ void __launch_bounds__(kSel2Block, 1) k_sel2_refine(const Sel2State* s2, const unsigned long long* __restrict__ cand, SelState* st) {
    __shared__ unsigned int hist[8192];
    __shared__ unsigned long long warp_tot[32];
    __shared__ int digit;
    __shared__ long long above_s;
    __shared__ unsigned int count_s;
    sel2_refine_body<false>((int64_t)s2->ncand, cand, nullptr, s2->bin, s2->remaining, s2->count_gt, hist, warp_tot, &digit, &above_s, &count_s, st);
}

// Whole selection in one CTA (m <= kSel2SmallMax): 15-bit histogram, pick, four refinement passes straight over g.
__global__ void __launch_bounds__(kSel2Block, 1) k_sel2_small(int64_t m, const double* __restrict__ g, long long k, SelState* st) {
    extern __shared__ unsigned int sh2[];   // [kSel2Bins]; its first 8192 words double as the refinement histogram
    __shared__ unsigned long long warp_tot[32];
    __shared__ int digit;
    __shared__ long long above_s;
    __shared__ unsigned int count_s;
    for (int i = threadIdx.x; i < kSel2Bins; i += kSel2Block) sh2[i] = 0u;
    __syncthreads();
    const int64_t mceil = ((m + 31) / 32) * 32;
    for (int64_t e = threadIdx.x; e < mceil; e += kSel2Block) {
        const bool ok = e < m;
        const unsigned int d = ok ? (unsigned int)(order_key(g[e]) >> 49) : 0u;
        sel2_hist_add(sh2, d, ok);
    }
    __syncthreads();
    sel2_find_from_top(sh2, kSel2Bins, k, warp_tot, &digit, &above_s, &count_s);
    const unsigned int bin = (unsigned int)digit;
    const long long rem = k - above_s, cgt = above_s;
    __syncthreads();
    sel2_refine_body<true>(m, nullptr, g, bin, rem, cgt, sh2, warp_tot, &digit, &above_s, &count_s, st);
}

// Selection mask + dual-bound term.  Block b owns the contiguous index range [b*chunk, (b+1)*chunk)
// so that ties at the k-th key can be ranked by index (lowest index first).
__global__ void __launch_bounds__(kBlock) k_sel_tie_count(int64_t m, int64_t chunk, const double* __restrict__ g,
                                                          const SelState* st, unsigned int* __restrict__ blockcnt) {
    __shared__ unsigned int cnt;
    if (threadIdx.x == 0) cnt = 0u;
    __syncthreads();
    const unsigned long long kth = st->prefix;
    const int64_t lo = (int64_t)blockIdx.x * chunk, hi = min(m, lo + chunk);
    unsigned int c = 0;
    for (int64_t e = lo + threadIdx.x; e < hi; e += blockDim.x) c += (order_key(g[e]) == kth) ? 1u : 0u;
    if (c) atomicAdd(&cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) blockcnt[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(kBlock) k_sel_tie_scan(int nblocks, unsigned int* __restrict__ blockcnt) {
    // exclusive scan by one thread: nblocks <= a few thousand, runs only when ties exist
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned int run = 0;
        for (int b = 0; b < nblocks; ++b) {
            unsigned int c = blockcnt[b];
            blockcnt[b] = run;
            run += c;
        }
    }
}

template <bool RANKED>
__global__ void __launch_bounds__(kBlock) k_sel_apply(int64_t m, int64_t chunk, const double* __restrict__ g,
                                                      const double* __restrict__ x, const SelState* st,
                                                      const unsigned int* __restrict__ blockoff,
                                                      uint8_t* __restrict__ sel, LzScalars* sc, ReduceWS ws) {
    __shared__ double sm[2 * kWarpsPerBlock];
    __shared__ int flag;
    __shared__ unsigned int run;       // ties seen so far in this block's range
    __shared__ unsigned int wcnt[kWarpsPerBlock];
    const unsigned long long kth = st->prefix;
    const long long need = st->remaining;   // ties to take, lowest index first
    const bool none = (st->mask == 0ull);   // k == 0: nothing selected
    const int64_t lo = (int64_t)blockIdx.x * chunk, hi = min(m, lo + chunk);
    if (threadIdx.x == 0) run = RANKED ? blockoff[blockIdx.x] : 0u;
    __syncthreads();
    double r0 = 0.0, r1 = 0.0;
    for (int64_t base = lo; base < hi; base += blockDim.x) {
        const int64_t e = base + threadIdx.x;
        bool in = e < hi;
        double ge = in ? g[e] : 0.0;
        unsigned long long key = in ? order_key(ge) : 0ull;
        bool take = in && !none && key > kth;
        bool tie = in && !none && key == kth;
        if (RANKED) {
            unsigned int bal = __ballot_sync(0xffffffffu, tie);
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            if (lane == 0) wcnt[warp] = __popc(bal);
            __syncthreads();
            unsigned int before = run;
            for (int w = 0; w < warp; ++w) before += wcnt[w];
            unsigned int rank = before + __popc(bal & ((1u << lane) - 1u));
            if (tie && (long long)rank < need) take = true;
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned int tot = 0;
                for (int w = 0; w < kWarpsPerBlock; ++w) tot += wcnt[w];
                run += tot;
            }
            __syncthreads();
        } else {
            take = take || tie;
        }
        if (in) {
            sel[e] = take ? 1 : 0;
            double sx = (take ? 1.0 : 0.0) - x[e];
            r0 = fma(ge, sx, r0);
            r1 += take ? 1.0 : 0.0;
        }
    }
    double r[2] = {r0, r1};
    if (grid_reduce<2>(r, ws, sm, &flag)) {
        sc->gs_minus_x = r[0];
        sc->nsel = (int64_t)(r[1] + 0.5);
    }
}

// ---- FW update: x <- x + gamma (s - x)  (frankwolfe.py:76), candidate edge weights refreshed -------
// The three roundings (s - x, gamma * (.), x + (.)) are kept separate, as numpy evaluates them.
// (x_out may be x itself, or a second buffer: the pipelined Frank-Wolfe loop keeps the iterate it has not yet accepted)
__global__ void __launch_bounds__(kBlock) k_fw_update(int64_t m, double gamma, const uint8_t* __restrict__ sel,
                                                      const double* __restrict__ kappa, double tol,
                                                      const double* x, double* x_out, double* __restrict__ ew_cand) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (int64_t)gridDim.x * blockDim.x) {
        double xe = x[e];
        double s = sel[e] ? 1.0 : 0.0;
        double xn = __dadd_rn(xe, __dmul_rn(gamma, __dsub_rn(s, xe)));
        x_out[e] = xn;
        ew_cand[e] = (xn > tol) ? xn * kappa[e] : 0.0;
    }
}

// ---- tie-broken nearest rounding (rounding.py:30-42) -----------------------------------------------------
// t = numpy.round(w, decimals): rint(w * 10^d) / 10^d, each operation rounded separately like numpy's C loop.
__global__ void __launch_bounds__(kBlock) k_round_decimals(int64_t m, const double* __restrict__ w, double p10,
                                                           double* __restrict__ t) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (int64_t)gridDim.x * blockDim.x)
        t[e] = __ddiv_rn(rint(__dmul_rn(w[e], p10)), p10);
}

// second key of the lexicographic order: the edge weight where the first key ties with the k-th value, -inf elsewhere
__global__ void __launch_bounds__(kBlock) k_tie_keys(int64_t m, const double* __restrict__ t, const double* __restrict__ kappa,
                                                     const SelState* st, double* __restrict__ out) {
    const unsigned long long kth = st->prefix;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = (order_key(t[e]) == kth) ? kappa[e] : -INFINITY;
}

// final mask: first key above the k-th value, or selected by the second-key pass
__global__ void __launch_bounds__(kBlock) k_round_merge(int64_t m, const double* __restrict__ t, const SelState* st1,
                                                        const uint8_t* __restrict__ sel2, double* __restrict__ out) {
    const unsigned long long kth = st1->prefix;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = (order_key(t[e]) > kth || sel2[e]) ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(kBlock) k_mask_to_double(int64_t m, const uint8_t* __restrict__ sel, double* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < m; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = sel[e] ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(kBlock) k_fill(int64_t count, double value, double* __restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = value;
}

// ---- measurement: L2 -> SM read bandwidth (the bound of the Lanczos kernels: their matrix is L2-resident) -------------
// Every CTA streams the whole buffer (16-byte loads that bypass L1) `reps` times, starting at a CTA-dependent offset so
// that the CTAs do not walk the slices in lock step.  bytes moved = gridDim.x * reps * bytes.
__global__ void __launch_bounds__(1024, 1) k_l2_read(const double2* __restrict__ buf, int64_t n16, int reps, double* __restrict__ sink) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const int64_t start = ((int64_t)blockIdx.x * 9973 * 1024) % n16;
    for (int r = 0; r < reps; ++r) {
        for (int64_t i0 = threadIdx.x; i0 < n16; i0 += 4 * 1024) {
            double2 v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                int64_t i = i0 + q * 1024 + start;
                if (i >= n16) i -= n16;
                v[q] = (i0 + q * 1024 < n16) ? __ldcg(buf + i) : make_double2(0.0, 0.0);
            }
            a0 += v[0].x + v[0].y; a1 += v[1].x + v[1].y; a2 += v[2].x + v[2].y; a3 += v[3].x + v[3].y;
        }
    }
    const double t = (a0 + a1) + (a2 + a3);
    if (t == 1.2345e-300) sink[0] = t;   // never true: keeps the loads alive
}

// coefficients not yet produced read as NaN; Rayleigh-Ritz state cleared (engines without an init kernel of their own for this)
__global__ void __launch_bounds__(kBlock) k_rr_reset(double* __restrict__ alpha, double* __restrict__ beta, int ncoef, RrOut* rr_out, int* dev_stop) {
    const double nan1 = __longlong_as_double(0x7ff8000000000001ll);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncoef; i += gridDim.x * blockDim.x) {
        alpha[i] = nan1;
        beta[i] = nan1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        rr_out->status = 0; rr_out->k = 0; rr_out->checks = 0; rr_out->phases = 0;
        rr_out->theta = 0.0; rr_out->est = 0.0; rr_out->target = 0.0;
        rr_out->cyc_wait = 0; rr_out->cyc_compute = 0; rr_out->lag = 0; rr_out->rounds = 0;
        *dev_stop = 0;
    }
}

__global__ void k_clear_lp_scalars(LzScalars* sc) {
    sc->gs_minus_x = 0.0;
    sc->nsel = 0;
}

}  // namespace macb
