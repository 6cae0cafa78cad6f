// C ABI of libbunmpc.so (include/bunmpc.h): solver handle, staging buffers, kernel dispatch.
// Host orchestration only; the arithmetic is in kernels.cuh.  No CPU fallback anywhere: if a CUDA call
// fails the entry point returns BUNMPC_ERR_CUDA and bunmpc_last_error() says why.
#include "../../include/bunmpc.h"
#include "kernels.cuh"
#include "solve_inst.hpp"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

using namespace bunmpc;

static thread_local std::string g_err;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            return BUNMPC_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

static int fail(int code, const std::string &msg) { g_err = msg; return code; }

struct bunmpc_solver {
    int device = 0, n = 0, e = 0, nx = 0, nf = 0, max_batch = 0, num_sms = 0;
    cudaStream_t stream = nullptr;
    unsigned int *job_counters = nullptr;   // [2] fresh-instance counters of a multi-GPU job (bunmpc_set_job_counter), device or peer memory
    int job_owner = 0; unsigned long long job_step = 0;
    PeerOut *peers = nullptr; int n_peers = 0;   // [kMaxPeers] result rows of the job's other GPUs (bunmpc_set_peer_results)
    unsigned int *work_counter = nullptr;   // [2 + 2 kParkQueues]: next fresh instance, finished instances, tail / head of each queue of parked instances
    int *queue = nullptr; double *sl_d = nullptr; int *sl_i = nullptr; long long *sl_c = nullptr;   // time slicing
    double *coef = nullptr;          // device, [coef_len]
    int coef_len = 0;
    // staging (device), sized for max_batch
    double *st_in = nullptr;         // compact / expanded inputs copied from the host
    size_t st_in_doubles = 0;
    double *ex = nullptr;            // expanded Qx,qx,lbx,ubx,Qf,qf
    double *out_d = nullptr;         // X,F,P,L,viol,(hist)
    int *out_i = nullptr;            // iters, status
    long long *out_c = nullptr;      // cycles
    double *mats = nullptr;          // scratch for bunmpc_centroidal_mats_host
    double *hist = nullptr;          // [max_batch][hist_cols] viol_hist staging of the host entry points (grown on demand)
    int hist_cols = 0;
    long long launches = 0;
    int nthreads = 0, smem_bytes = 0;
    int ctas_per_sm[3] = {0, 0, 0};  // per arith
};

static size_t smem_bytes_for(int n, int e, int max_inner, int nthreads)
{
    return (size_t)make_layout(n, e, max_inner, nthreads / 32).total * sizeof(double);
}

static const int kMaxSmemBytes = 227 * 1024 - 256;  // dynamic + the kernel's static shared memory (~100 bytes) must fit the 227 KB opt-in limit

static void free_solver(bunmpc_solver *s)
{
    if (!s) return;
    cudaFree(s->queue); cudaFree(s->sl_d); cudaFree(s->sl_i); cudaFree(s->sl_c);
    cudaFree(s->peers);
    cudaFree(s->work_counter); cudaFree(s->coef); cudaFree(s->st_in); cudaFree(s->ex);
    cudaFree(s->out_d); cudaFree(s->out_i); cudaFree(s->out_c); cudaFree(s->mats); cudaFree(s->hist);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

// like CK, but releases a half-built solver first
#define CKS(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            free_solver(s);                                                                        \
            return BUNMPC_ERR_CUDA;                                                                \
        }                                                                                          \
    } while (0)

extern "C" {

int bunmpc_version(void) { return BUNMPC_VERSION; }
const char *bunmpc_last_error(void) { return g_err.c_str(); }

void bunmpc_default_params(bunmpc_params *p)
{
    p->max_outer = 100; p->max_inner = 150; p->tol = 1e-5; p->exit_tol = 1e-3; p->beta = 1.5; p->mu = 1.0;
    p->arith = BUNMPC_ARITH_STRICT;
    p->slice_outer = 0;
}

// ---- one fresh-instance counter for all ranks of a multi-GPU job: a device allocation of the owner, exported as a CUDA
// IPC handle, mapped by the other ranks' processes (peer memory over NVLink) ----
int bunmpc_job_counter_create(int device, void **ptr, unsigned char handle[64])
{
    if (!ptr || !handle) return fail(BUNMPC_ERR_ARG, "job_counter_create: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t");
    CK(cudaSetDevice(device));
    void *p = nullptr;
    CK(cudaMalloc(&p, 256));
    CK(cudaMemset(p, 0, 256));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(BUNMPC_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    memcpy(handle, &h, 64);
    *ptr = p;
    return BUNMPC_OK;
}

int bunmpc_job_counter_open(int device, const unsigned char handle[64], void **ptr)
{
    if (!ptr || !handle) return fail(BUNMPC_ERR_ARG, "job_counter_open: null argument");
    CK(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(BUNMPC_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
    *ptr = p;
    return BUNMPC_OK;
}

int bunmpc_job_counter_release(void *ptr, int owner)
{
    if (!ptr) return BUNMPC_OK;
    cudaError_t e = owner ? cudaFree(ptr) : cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) return fail(BUNMPC_ERR_CUDA, std::string("job_counter_release: ") + cudaGetErrorString(e));
    return BUNMPC_OK;
}

int bunmpc_set_job_counter(bunmpc_solver *s, void *counters, int owner)
{
    if (!s) return fail(BUNMPC_ERR_ARG, "set_job_counter: null solver");
    s->job_counters = (unsigned int *)counters; s->job_owner = owner ? 1 : 0; s->job_step = 0;
    return BUNMPC_OK;
}

// ---- result buffers that the other GPUs of a job can store into: a device allocation exported as a CUDA IPC handle ----
int bunmpc_peer_buffer_create(int device, unsigned long long bytes, void **ptr, unsigned char handle[64])
{
    if (!ptr || !handle || bytes == 0) return fail(BUNMPC_ERR_ARG, "peer_buffer_create: bad argument");
    CK(cudaSetDevice(device));
    void *p = nullptr;
    CK(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail(BUNMPC_ERR_CUDA, std::string("peer_buffer_create: ") + cudaGetErrorString(e)); }
    memcpy(handle, &h, 64);
    *ptr = p;
    return BUNMPC_OK;
}

int bunmpc_peer_buffer_open(int device, const unsigned char handle[64], void **ptr)
{
    return bunmpc_job_counter_open(device, handle, ptr);
}

int bunmpc_peer_buffer_release(void *ptr, int owner)
{
    return bunmpc_job_counter_release(ptr, owner);
}

int bunmpc_set_peer_results(bunmpc_solver *s, int n_peers, const bunmpc_solution *peers)
{
    if (!s || n_peers < 0 || n_peers > kMaxPeers || (n_peers > 0 && !peers))
        return fail(BUNMPC_ERR_ARG, "set_peer_results: bad argument");
    CK(cudaSetDevice(s->device));
    PeerOut h[kMaxPeers];
    for (int g = 0; g < n_peers; ++g) {
        const bunmpc_solution &q = peers[g];
        if (!q.X || !q.F || !q.L || !q.viol || !q.iters || !q.status)
            return fail(BUNMPC_ERR_ARG, "set_peer_results: X, F, L, viol, iters and status are required for every peer");
        h[g] = PeerOut{q.X, q.F, q.L, q.viol, q.iters, q.status};
    }
    if (n_peers > 0) {
        if (!s->peers) CK(cudaMalloc(&s->peers, sizeof(PeerOut) * kMaxPeers));
        CK(cudaMemcpy(s->peers, h, sizeof(PeerOut) * n_peers, cudaMemcpyHostToDevice));
    }
    s->n_peers = n_peers;
    return BUNMPC_OK;
}

void *bunmpc_host_alloc(unsigned long long bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void bunmpc_host_free(void *p) { if (p) cudaFreeHost(p); }

static const int kMaxInnerTable = 4096;
static const int kQueuePerInstance = 16;   // an instance is parked at most 15 times per solve

int bunmpc_create(bunmpc_solver **out, int device, int n_col, int n_eff, int max_batch)
{
    if (!out || n_col < 1 || max_batch < 1) return fail(BUNMPC_ERR_ARG, "bunmpc_create: bad argument");
    if (n_eff != 4) return fail(BUNMPC_ERR_UNSUPPORTED, "bunmpc_create: kernels are built for n_eff == 4");
    const int n = n_col, e = n_eff, nx = 9 * (n + 1), nf = 3 * e * n;
    int nthreads = solve_threads(n, e);                   // e*n force threads / 3(n+1) state threads, one CTA per instance
    if (nthreads == 0) return fail(BUNMPC_ERR_UNSUPPORTED, "bunmpc_create: n_col too large for one CTA per instance");
    // a horizon whose pipelined layout does not fit into the shared memory of one SM runs in the lean layout of the big
    // CTAs (kernels.cuh: big_cta)
    if (nthreads < 512 && smem_bytes_for(n, e, 150, nthreads) > (size_t)kMaxSmemBytes) nthreads = 512;
    if (smem_bytes_for(n, e, 150, nthreads) > (size_t)kMaxSmemBytes)
        return fail(BUNMPC_ERR_UNSUPPORTED, "bunmpc_create: n_col too large for the shared memory of one SM");
    CK(cudaSetDevice(device));
    bunmpc_solver *s = new bunmpc_solver();
    s->device = device; s->n = n; s->e = e; s->nx = nx; s->nf = nf; s->max_batch = max_batch;
    s->nthreads = nthreads;
    cudaDeviceProp prop;
    CKS(cudaGetDeviceProperties(&prop, device));
    s->num_sms = prop.multiProcessorCount;
    CKS(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));

    // FISTA momentum coefficients (t_k - 1)/t_{k+1} with t_{k+1} = 1 + sqrt(1 + 4 t_k^2)/2 (fista.cpp:34-35, sic);
    // the sequence does not depend on the data, so it is tabulated once (IEEE sqrt and / are exact on both sides).
    {
        std::vector<double> c(kMaxInnerTable);
        volatile double t_k = 1.0;
        for (int i = 0; i < kMaxInnerTable; ++i) {
            volatile double sq = 4 * t_k * t_k;
            volatile double t_k_1 = 1.0 + std::sqrt(1 + sq) / 2.0;
            c[i] = (t_k - 1) / t_k_1;
            t_k = t_k_1;
        }
        CKS(cudaMalloc(&s->coef, sizeof(double) * kMaxInnerTable));
        CKS(cudaMemcpy(s->coef, c.data(), sizeof(double) * kMaxInnerTable, cudaMemcpyHostToDevice));
        s->coef_len = kMaxInnerTable;
    }
    CKS(cudaMalloc(&s->work_counter, kWorkCounters * sizeof(unsigned int)));
    CKS(cudaMalloc(&s->queue, sizeof(int) * kParkQueues * kQueuePerInstance * (size_t)max_batch));
    CKS(cudaMalloc(&s->sl_d, sizeof(double) * (size_t)max_batch * (2 * (size_t)nx + nf + 2)));
    CKS(cudaMalloc(&s->sl_i, sizeof(int) * 8 * (size_t)max_batch));
    CKS(cudaMalloc(&s->sl_c, sizeof(long long) * (size_t)max_batch));

    // staging buffers
    const size_t B = (size_t)max_batch;
    const size_t in_doubles = B * (2 + 9 + 4 * (size_t)e * n + n + 2 + 2 * (size_t)nx + nf     // m,rho,x_init,cnt,dt,L0,X0,P0,F0
                                   + 4 * (size_t)nx + 2 * (size_t)nf + 6 * (size_t)n + 18) + 64;  // costs/bounds in either form
    s->st_in_doubles = in_doubles;
    CKS(cudaMalloc(&s->st_in, sizeof(double) * in_doubles));
    CKS(cudaMalloc(&s->ex, sizeof(double) * B * (4 * (size_t)nx + 2 * (size_t)nf)));
    CKS(cudaMalloc(&s->out_d, sizeof(double) * B * (2 * (size_t)nx + nf + 3)));
    CKS(cudaMalloc(&s->out_i, sizeof(int) * B * 6));
    CKS(cudaMalloc(&s->out_c, sizeof(long long) * B));
    CKS(cudaMalloc(&s->mats, sizeof(double) * ((size_t)nx * nf + (size_t)nx * nx + 2 * (size_t)nx + 4 * (size_t)e * n + n + nx + nf + 9)));

    // opt in to the shared memory the kernel needs and record occupancy
    s->smem_bytes = (int)smem_bytes_for(n, e, 150, nthreads);
    for (int arith = 0; arith < 3; ++arith) {
        solve_fn fn = solve_pick(nthreads, arith);
        CKS(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes));
        CKS(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        int nb = 0;
        CKS(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, nthreads, s->smem_bytes));
        s->ctas_per_sm[arith] = nb;
    }
    *out = s;
    return BUNMPC_OK;
}

void bunmpc_destroy(bunmpc_solver *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    free_solver(s);
}

long long bunmpc_launch_count(const bunmpc_solver *s) { return s ? s->launches : 0; }

int bunmpc_kernel_info(const bunmpc_solver *s, int *ctas_per_sm, int *threads, int *smem_bytes, int *num_sms)
{
    if (!s) return fail(BUNMPC_ERR_ARG, "null solver");
    if (ctas_per_sm) *ctas_per_sm = s->ctas_per_sm[0];
    if (threads) *threads = s->nthreads;
    if (smem_bytes) *smem_bytes = s->smem_bytes;
    if (num_sms) *num_sms = s->num_sms;
    return BUNMPC_OK;
}

static In mk(const bunmpc_in &f) { return In{f.ptr, f.batch_stride}; }

static int check_params(const bunmpc_solver *s, const bunmpc_params *prm)
{
    if (!prm) return fail(BUNMPC_ERR_ARG, "null params");
    if (prm->max_outer < 0 || prm->max_inner < 1 || prm->max_inner > s->coef_len)
        return fail(BUNMPC_ERR_ARG, "max_inner/max_outer out of range");
    if (prm->arith != BUNMPC_ARITH_STRICT && prm->arith != BUNMPC_ARITH_FMA && prm->arith != BUNMPC_ARITH_MIXED)
        return fail(BUNMPC_ERR_UNSUPPORTED, "unknown arith mode");
    // a rejected step multiplies L by beta until it is accepted (fista.cpp:19): beta <= 1 would spin forever inside a
    // persistent kernel; the tolerances only have to be comparable
    if (!(prm->beta > 1.0) || !std::isfinite(prm->beta)) return fail(BUNMPC_ERR_ARG, "beta must be finite and > 1");
    if (std::isnan(prm->tol) || std::isnan(prm->exit_tol) || std::isnan(prm->mu)) return fail(BUNMPC_ERR_ARG, "tol / exit_tol / mu is NaN");
    return BUNMPC_OK;
}

int bunmpc_expand_device(bunmpc_solver *s, const bunmpc_compact_problem *p, double *Qx, double *qx, double *Qf,
                         double *qf, double *lbx, double *ubx, void *stream)
{
    if (!s || !p || !Qx || !qx || !Qf || !qf || !lbx || !ubx) return fail(BUNMPC_ERR_ARG, "expand: null argument");
    if (p->batch < 1) return fail(BUNMPC_ERR_ARG, "expand: batch < 1");
    CK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;   // NULL is CUDA's default stream
    ExpandArgs a;
    a.B = p->batch; a.n = s->n; a.e = s->e; a.nx = s->nx; a.nf = s->nf;
    a.cnt_plan = mk(p->cnt_plan); a.W_X = mk(p->W_X); a.W_X_ter = mk(p->W_X_ter); a.X_nom = mk(p->X_nom);
    a.X_ter = mk(p->X_ter); a.W_F = mk(p->W_F); a.bounds = mk(p->bounds);
    a.Qx = Qx; a.qx = qx; a.Qf = Qf; a.qf = qf; a.lbx = lbx; a.ubx = ubx;
    const long long total = (long long)a.B * (a.nx > a.nf ? a.nx : a.nf);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)s->num_sms * 8;
    if (blocks > cap) blocks = cap;
    expand_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
    s->launches++;
    CK(cudaGetLastError());
    return BUNMPC_OK;
}

int bunmpc_solve_expanded_device(bunmpc_solver *s, const bunmpc_expanded_problem *p, const bunmpc_params *prm,
                                 const bunmpc_solution *out, void *stream)
{
    if (!s || !p || !out) return fail(BUNMPC_ERR_ARG, "solve: null argument");
    int rc = check_params(s, prm);
    if (rc) return rc;
    // the time-slicing scratch (queue, parked states) is sized by max_batch
    if (p->batch < 1 || p->batch > s->max_batch) return fail(BUNMPC_ERR_ARG, "solve: batch outside [1, max_batch]");
    if (!p->m.ptr || !p->rho.ptr || !p->x_init.ptr || !p->cnt_plan.ptr || !p->dt.ptr || !p->Qx.ptr || !p->qx.ptr ||
        !p->Qf.ptr || !p->qf.ptr || !p->lbx.ptr || !p->ubx.ptr || !p->L0.ptr)
        return fail(BUNMPC_ERR_ARG, "solve: null input field");
    CK(cudaSetDevice(s->device));
    cudaStream_t st = (cudaStream_t)stream;   // NULL is CUDA's default stream
    SolveArgs a;
    a.B = p->batch; a.n = s->n;
    a.m = mk(p->m); a.rho = mk(p->rho); a.x_init = mk(p->x_init); a.cnt_plan = mk(p->cnt_plan); a.dt = mk(p->dt);
    a.Qx = mk(p->Qx); a.qx = mk(p->qx); a.Qf = mk(p->Qf); a.qf = mk(p->qf); a.lbx = mk(p->lbx); a.ubx = mk(p->ubx);
    a.L0 = mk(p->L0); a.X0 = mk(p->X0); a.F0 = mk(p->F0); a.P0 = mk(p->P0);
    a.X = out->X; a.F = out->F; a.P = out->P; a.L = out->L; a.viol = out->viol; a.viol_hist = out->viol_hist;
    a.iters = out->iters; a.status = out->status; a.cycles = out->cycles; a.prof = nullptr;
#ifdef BUNMPC_PHASE_PROF
    a.prof = reinterpret_cast<long long *>(out->viol_hist);   // profiling build: the viol_hist buffer carries [B][16] counters
    a.viol_hist = nullptr;
#endif
    a.max_outer = prm->max_outer; a.max_inner = prm->max_inner;
    a.tol = prm->tol; a.exit_tol = prm->exit_tol; a.beta = prm->beta; a.mu = prm->mu;
    a.coef = s->coef; a.work_counter = s->work_counter;
    a.S = make_layout(s->n, s->e, prm->max_inner, s->nthreads / 32);
    const int smem = (int)smem_bytes_for(s->n, s->e, prm->max_inner, s->nthreads);
    if (smem > kMaxSmemBytes) return fail(BUNMPC_ERR_UNSUPPORTED, "solve: shared memory need exceeds one SM");
    solve_fn fn = solve_pick(s->nthreads, prm->arith);
    int per_sm = s->ctas_per_sm[prm->arith];
    if (const char *ev = getenv("BUNMPC_CTAS")) {      // experiment: occupancy variant of the 128-thread kernel
        solve_fn alt = (s->nthreads == 128) ? solve_inst_x128(prm->arith, atoi(ev)) : nullptr;
        if (alt) {
            fn = alt;
            CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes));
            CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, s->nthreads, smem));
        }
    }
    else if (smem != s->smem_bytes) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, s->nthreads, smem));
    if (per_sm < 1) return fail(BUNMPC_ERR_UNSUPPORTED, "solve: kernel does not fit on an SM");
    long long grid = (long long)s->num_sms * per_sm;
    if (const char *ev = getenv("BUNMPC_MAX_CTAS")) {   // test knob: few resident CTAs, so that a handful of instances exercises the time slicing
        const long long cap = atoll(ev);
        if (cap >= 1 && cap < grid) grid = cap;
    }
    if (grid > a.B) grid = a.B;
    // time slicing: with more instances than resident CTAs, park an instance after a few outer iterations so that
    // the launch ends on a short slice instead of on the longest instance (iteration counts spread 4x)
    int slice = 0;
    if (a.B > grid) {
        const int min_slice = (prm->max_outer + kQueuePerInstance - 1) / kQueuePerInstance;   // <= 15 parks per instance
        slice = prm->slice_outer < 0 ? 0 : (prm->slice_outer > 0 ? prm->slice_outer : 8);
        if (slice > 0 && slice < min_slice) slice = min_slice;
    }
    a.slice_outer = slice; a.queue_cap = kQueuePerInstance * a.B;
    a.queue = s->queue; a.sl_d = s->sl_d; a.sl_i = s->sl_i; a.sl_c = s->sl_c;
    a.long_inner = 1000.f;      // scheduling heuristic (kernels.cuh, parking code; profiles/r02_sched_probe3.txt); never changes a result
    if (const char *ev = getenv("BUNMPC_LONG_INNER")) a.long_inner = (float)atof(ev);
    CK(cudaMemsetAsync(s->work_counter, 0, kWorkCounters * sizeof(unsigned int), st));
    a.peers = s->peers; a.n_peers = s->n_peers;
    a.job_counter = nullptr;
    if (s->job_counters) {
        // solve k of the job pulls from counter k & 1; the owner clears the other one for solve k + 1.  Between two solves
        // of a job every rank runs a collective (the exchange of the results), so that counter is idle now and stays idle
        // until this rank's kernel -- which follows the memset in stream order -- has finished
        a.job_counter = s->job_counters + (s->job_step & 1ull);
        if (s->job_owner) CK(cudaMemsetAsync(s->job_counters + ((s->job_step + 1ull) & 1ull), 0, sizeof(unsigned int), st));
        s->job_step++;
    }
    if (slice > 0) CK(cudaMemsetAsync(s->queue, 0xff, sizeof(int) * kParkQueues * (size_t)a.queue_cap, st));
    fn<<<(unsigned)grid, s->nthreads, smem, st>>>(a);
    s->launches++;
    CK(cudaGetLastError());
    return BUNMPC_OK;
}

int bunmpc_solve_compact_device(bunmpc_solver *s, const bunmpc_compact_problem *p, const bunmpc_params *prm,
                                const bunmpc_solution *out, void *stream)
{
    if (!s || !p || !out) return fail(BUNMPC_ERR_ARG, "solve: null argument");
    if (p->batch < 1 || p->batch > s->max_batch) return fail(BUNMPC_ERR_ARG, "solve: batch outside [1, max_batch]");
    const size_t B = (size_t)p->batch, nx = (size_t)s->nx, nf = (size_t)s->nf;
    double *Qx = s->ex, *qx = Qx + B * nx, *lbx = qx + B * nx, *ubx = lbx + B * nx, *Qf = ubx + B * nx, *qf = Qf + B * nf;
    int rc = bunmpc_expand_device(s, p, Qx, qx, Qf, qf, lbx, ubx, stream);
    if (rc) return rc;
    bunmpc_expanded_problem q;
    q.batch = p->batch;
    q.m = p->m; q.rho = p->rho; q.x_init = p->x_init; q.cnt_plan = p->cnt_plan; q.dt = p->dt;
    q.Qx = {Qx, (long long)nx}; q.qx = {qx, (long long)nx}; q.Qf = {Qf, (long long)nf}; q.qf = {qf, (long long)nf};
    q.lbx = {lbx, (long long)nx}; q.ubx = {ubx, (long long)nx};
    q.L0 = p->L0; q.X0 = p->X0; q.F0 = p->F0; q.P0 = p->P0;
    return bunmpc_solve_expanded_device(s, &q, prm, out, stream);
}

int bunmpc_build_problem_device(bunmpc_solver *s, const bunmpc_gait *g, const bunmpc_states *st, double *x_init,
                                double *cnt_plan, double *dt, double *X_nom, double *X_ter, double *W_X,
                                double *W_X_ter, double *W_F, double *rho, void *stream)
{
    if (!s || !g || !st || !x_init || !cnt_plan || !dt || !X_nom || !X_ter) return fail(BUNMPC_ERR_ARG, "build: null argument");
    if (st->batch < 1) return fail(BUNMPC_ERR_ARG, "build: batch < 1");
    if (!st->com.ptr || !st->vcom.ptr || !st->amom.ptr || !st->foot_pos.ptr || !st->t.ptr || !st->v_des.ptr ||
        !st->w_des.ptr || (!st->cs_yaw.ptr && !st->hip_xy.ptr)) return fail(BUNMPC_ERR_ARG, "build: null state field");
    if (st->scales.ptr && (!W_X || !W_X_ter || !W_F || !rho)) return fail(BUNMPC_ERR_ARG, "build: scales need W_X, W_X_ter, W_F, rho outputs");
    if (s->e != 4) return fail(BUNMPC_ERR_UNSUPPORTED, "build: quadrupeds only");
    CK(cudaSetDevice(s->device));
    BuildArgs a;
    a.B = st->batch; a.n = s->n;
    a.com = mk(st->com); a.vcom = mk(st->vcom); a.amom = mk(st->amom); a.foot_pos = mk(st->foot_pos); a.t = mk(st->t);
    a.v_des = mk(st->v_des); a.w_des = mk(st->w_des); a.cs_yaw = mk(st->cs_yaw); a.hip_xy = mk(st->hip_xy); a.amom_des = mk(st->amom_des);
    a.scales = mk(st->scales);
    a.x_init = x_init; a.cnt_plan = cnt_plan; a.dt = dt; a.X_nom = X_nom; a.X_ter = X_ter;
    a.W_X = W_X; a.W_X_ter = W_X_ter; a.W_F = W_F; a.rho = rho;
    static_assert(sizeof(GaitDev) == sizeof(bunmpc_gait), "bunmpc_gait layout");
    if (g->swing_rule != 0 && g->swing_rule != 1) return fail(BUNMPC_ERR_ARG, "build_problem: swing_rule must be 0 or 1");
    memcpy(&a.g, g, sizeof(GaitDev));
    build_problem_kernel<<<(a.B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
    s->launches++;
    CK(cudaGetLastError());
    return BUNMPC_OK;
}

int bunmpc_build_acyclic_device(bunmpc_solver *s, const bunmpc_acyclic_motion *m, int batch, const bunmpc_in *x_init,
                                const bunmpc_in *t, double *cnt_plan, double *dt, double *X_nom, double *X_ter,
                                double *bounds, void *stream)
{
    if (!s || !m || !x_init || !t || !x_init->ptr || !t->ptr || !cnt_plan || !dt || !X_nom || !X_ter || !bounds)
        return fail(BUNMPC_ERR_ARG, "build_acyclic: null argument");
    if (!m->dt_arr || !m->cnt_plan || !m->X_nom || !m->bounds || !m->X_ter || m->n_cnt < 1 || m->n_nom < 1 || m->n_box < 1)
        return fail(BUNMPC_ERR_ARG, "build_acyclic: incomplete motion tables");
    if (batch < 1) return fail(BUNMPC_ERR_ARG, "build_acyclic: batch < 1");
    if (s->e != 4) return fail(BUNMPC_ERR_UNSUPPORTED, "build_acyclic: four end effectors");
    CK(cudaSetDevice(s->device));
    AcyclicArgs a;
    a.B = batch; a.n = s->n; a.n_cnt = m->n_cnt; a.n_nom = m->n_nom; a.n_box = m->n_box;
    a.dt_arr = m->dt_arr; a.cnt = m->cnt_plan; a.nom = m->X_nom; a.box = m->bounds; a.X_ter_rec = m->X_ter; a.t0 = m->t0;
    a.x_init = mk(*x_init); a.t = mk(*t);
    a.cnt_plan = cnt_plan; a.dt = dt; a.X_nom = X_nom; a.X_ter = X_ter; a.bounds = bounds;
    build_acyclic_kernel<<<(batch + 127) / 128, 128, 0, (cudaStream_t)stream>>>(a);
    s->launches++;
    CK(cudaGetLastError());
    return BUNMPC_OK;
}

int bunmpc_goal_stats_device(bunmpc_solver *s, int batch, const bunmpc_in *goals, const bunmpc_in *errors, double *out17,
                             void *stream)
{
    if (!s || !goals || !errors || !goals->ptr || !errors->ptr || !out17 || batch < 1)
        return fail(BUNMPC_ERR_ARG, "goal_stats: bad argument");
    CK(cudaSetDevice(s->device));
    goal_stats_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(batch, mk(*goals), mk(*errors), out17);
    s->launches++;
    CK(cudaGetLastError());
    return BUNMPC_OK;
}

// ---- host-pointer path: stage fields into st_in, run, copy results back ----
struct Stager {
    bunmpc_solver *s;
    double *cur;
    int B;
    cudaError_t err = cudaSuccess;
    bunmpc_in put(const bunmpc_in &h, size_t width)
    {
        bunmpc_in d{nullptr, 0};
        if (!h.ptr || err != cudaSuccess) return d;
        d.ptr = cur;
        if (h.batch_stride == 0 || B == 1) {
            err = cudaMemcpyAsync(cur, h.ptr, width * sizeof(double), cudaMemcpyHostToDevice, s->stream);
            d.batch_stride = 0;
            cur += width;
        } else if ((size_t)h.batch_stride == width) {
            err = cudaMemcpyAsync(cur, h.ptr, (size_t)B * width * sizeof(double), cudaMemcpyHostToDevice, s->stream);
            d.batch_stride = (long long)width;
            cur += (size_t)B * width;
        } else {
            err = cudaMemcpy2DAsync(cur, width * sizeof(double), h.ptr, (size_t)h.batch_stride * sizeof(double),
                                    width * sizeof(double), (size_t)B, cudaMemcpyHostToDevice, s->stream);
            d.batch_stride = (long long)width;
            cur += (size_t)B * width;
        }
        cur += (size_t)(cur - s->st_in) & 1;   // keep 16-byte alignment
        return d;
    }
};

static int copy_out(bunmpc_solver *s, int B, const bunmpc_solution &dev, const bunmpc_solution *out, int max_outer)
{
    const size_t nx = (size_t)s->nx, nf = (size_t)s->nf, Bs = (size_t)B;
    if (out->X) CK(cudaMemcpyAsync(out->X, dev.X, Bs * nx * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (out->F) CK(cudaMemcpyAsync(out->F, dev.F, Bs * nf * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (out->P) CK(cudaMemcpyAsync(out->P, dev.P, Bs * nx * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (out->L) CK(cudaMemcpyAsync(out->L, dev.L, Bs * 2 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (out->viol) CK(cudaMemcpyAsync(out->viol, dev.viol, Bs * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    if (out->iters) CK(cudaMemcpyAsync(out->iters, dev.iters, Bs * 5 * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    if (out->status) CK(cudaMemcpyAsync(out->status, dev.status, Bs * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    if (out->cycles) CK(cudaMemcpyAsync(out->cycles, dev.cycles, Bs * sizeof(long long), cudaMemcpyDeviceToHost, s->stream));
    if (out->viol_hist)
        CK(cudaMemcpyAsync(out->viol_hist, dev.viol_hist, Bs * (size_t)max_outer * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return BUNMPC_OK;
}

static int dev_solution(bunmpc_solver *s, int B, const bunmpc_solution *want, int max_outer, bunmpc_solution *dev,
                        double **hist_alloc)
{
    const size_t nx = (size_t)s->nx, nf = (size_t)s->nf, Bs = (size_t)B;
    dev->X = s->out_d; dev->F = dev->X + Bs * nx; dev->P = dev->F + Bs * nf; dev->L = dev->P + Bs * nx;
    dev->viol = dev->L + 2 * Bs;
    dev->iters = s->out_i; dev->status = s->out_i + 5 * Bs;
    dev->cycles = want->cycles ? s->out_c : nullptr;
    dev->viol_hist = nullptr;
    *hist_alloc = nullptr;
    if (want->viol_hist) {
        // staging for return_dyn_viol_hist: sized [max_batch][max_outer] the first time a history is asked for (and
        // again only if a later call raises max_outer); solves that do not collect statistics never allocate
        if (s->hist_cols < max_outer) {
            CK(cudaStreamSynchronize(s->stream));
            cudaFree(s->hist); s->hist = nullptr; s->hist_cols = 0;
            CK(cudaMalloc(&s->hist, (size_t)s->max_batch * (size_t)max_outer * sizeof(double)));
            s->hist_cols = max_outer;
        }
        dev->viol_hist = s->hist;
    }
    return BUNMPC_OK;
}

int bunmpc_solve_compact_host(bunmpc_solver *s, const bunmpc_compact_problem *p, const bunmpc_params *prm,
                              const bunmpc_solution *out)
{
    if (!s || !p || !out) return fail(BUNMPC_ERR_ARG, "solve: null argument");
    int rc = check_params(s, prm);
    if (rc) return rc;
    if (p->batch < 1 || p->batch > s->max_batch) return fail(BUNMPC_ERR_ARG, "solve: batch outside [1, max_batch]");
    CK(cudaSetDevice(s->device));
    const size_t n = (size_t)s->n, e = (size_t)s->e, nx = (size_t)s->nx, nf = (size_t)s->nf;
    Stager st{s, s->st_in, p->batch};
    bunmpc_compact_problem d;
    d.batch = p->batch;
    d.m = st.put(p->m, 1); d.rho = st.put(p->rho, 1); d.x_init = st.put(p->x_init, 9);
    d.cnt_plan = st.put(p->cnt_plan, 4 * e * n); d.dt = st.put(p->dt, n);
    d.W_X = st.put(p->W_X, 9 * n); d.W_X_ter = st.put(p->W_X_ter, 9); d.X_nom = st.put(p->X_nom, 9 * n);
    d.X_ter = st.put(p->X_ter, 9); d.W_F = st.put(p->W_F, nf); d.bounds = st.put(p->bounds, 6 * n);
    d.L0 = st.put(p->L0, 2); d.X0 = st.put(p->X0, nx); d.F0 = st.put(p->F0, nf); d.P0 = st.put(p->P0, nx);
    CK(st.err);
    bunmpc_solution dev;
    double *hist = nullptr;
    rc = dev_solution(s, p->batch, out, prm->max_outer, &dev, &hist);
    if (rc) return rc;
    rc = bunmpc_solve_compact_device(s, &d, prm, &dev, s->stream);
    if (!rc) rc = copy_out(s, p->batch, dev, out, prm->max_outer);
    (void)hist;
    return rc;
}

int bunmpc_solve_expanded_host(bunmpc_solver *s, const bunmpc_expanded_problem *p, const bunmpc_params *prm,
                               const bunmpc_solution *out)
{
    if (!s || !p || !out) return fail(BUNMPC_ERR_ARG, "solve: null argument");
    int rc = check_params(s, prm);
    if (rc) return rc;
    if (p->batch < 1 || p->batch > s->max_batch) return fail(BUNMPC_ERR_ARG, "solve: batch outside [1, max_batch]");
    CK(cudaSetDevice(s->device));
    const size_t n = (size_t)s->n, e = (size_t)s->e, nx = (size_t)s->nx, nf = (size_t)s->nf;
    Stager st{s, s->st_in, p->batch};
    bunmpc_expanded_problem d;
    d.batch = p->batch;
    d.m = st.put(p->m, 1); d.rho = st.put(p->rho, 1); d.x_init = st.put(p->x_init, 9);
    d.cnt_plan = st.put(p->cnt_plan, 4 * e * n); d.dt = st.put(p->dt, n);
    d.Qx = st.put(p->Qx, nx); d.qx = st.put(p->qx, nx); d.Qf = st.put(p->Qf, nf); d.qf = st.put(p->qf, nf);
    d.lbx = st.put(p->lbx, nx); d.ubx = st.put(p->ubx, nx);
    d.L0 = st.put(p->L0, 2); d.X0 = st.put(p->X0, nx); d.F0 = st.put(p->F0, nf); d.P0 = st.put(p->P0, nx);
    CK(st.err);
    bunmpc_solution dev;
    double *hist = nullptr;
    rc = dev_solution(s, p->batch, out, prm->max_outer, &dev, &hist);
    if (rc) return rc;
    rc = bunmpc_solve_expanded_device(s, &d, prm, &dev, s->stream);
    if (!rc) rc = copy_out(s, p->batch, dev, out, prm->max_outer);
    (void)hist;
    return rc;
}

int bunmpc_centroidal_mats_host(bunmpc_solver *s, double m, const double *cnt_plan, const double *dt, const double *X,
                                const double *F, const double *x_init, double *A_x, double *b_x, double *A_f,
                                double *b_f)
{
    if (!s || !cnt_plan || !dt) return fail(BUNMPC_ERR_ARG, "mats: null argument");
    if ((A_x || b_x) && !X) return fail(BUNMPC_ERR_ARG, "mats: X required for A_x/b_x");
    if ((A_f || b_f) && (!F || !x_init)) return fail(BUNMPC_ERR_ARG, "mats: F and x_init required for A_f/b_f");
    CK(cudaSetDevice(s->device));
    const size_t n = (size_t)s->n, e = (size_t)s->e, nx = (size_t)s->nx, nf = (size_t)s->nf;
    double *dAx = s->mats, *dAf = dAx + nx * nf, *dbx = dAf + nx * nx, *dbf = dbx + nx;
    double *dcnt = dbf + nx, *ddt = dcnt + 4 * e * n, *dX = ddt + n, *dF = dX + nx, *dxi = dF + nf;
    cudaStream_t st = s->stream;
    CK(cudaMemcpyAsync(dcnt, cnt_plan, 4 * e * n * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ddt, dt, n * sizeof(double), cudaMemcpyHostToDevice, st));
    if (X) CK(cudaMemcpyAsync(dX, X, nx * sizeof(double), cudaMemcpyHostToDevice, st));
    if (F) CK(cudaMemcpyAsync(dF, F, nf * sizeof(double), cudaMemcpyHostToDevice, st));
    if (x_init) CK(cudaMemcpyAsync(dxi, x_init, 9 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(dAx, 0, (nx * nf + nx * nx) * sizeof(double), st));
    dense_mats_kernel<4><<<4, 128, 0, st>>>((int)n, m, dcnt, ddt, X ? dX : nullptr, F ? dF : nullptr, dxi,
                                            (A_x ? dAx : nullptr), (b_x ? dbx : nullptr), (A_f ? dAf : nullptr),
                                            (b_f ? dbf : nullptr));
    s->launches++;
    CK(cudaGetLastError());
    if (A_x) CK(cudaMemcpyAsync(A_x, dAx, nx * nf * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (b_x) CK(cudaMemcpyAsync(b_x, dbx, nx * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (A_f) CK(cudaMemcpyAsync(A_f, dAf, nx * nx * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (b_f) CK(cudaMemcpyAsync(b_f, dbf, nx * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BUNMPC_OK;
}

int bunmpc_selftest_division(bunmpc_solver *s, long long n_pairs, unsigned long long seed, long long *mismatches)
{
    if (!s || !mismatches || n_pairs < 1) return fail(BUNMPC_ERR_ARG, "selftest: bad argument");
    CK(cudaSetDevice(s->device));
    unsigned long long *d = reinterpret_cast<unsigned long long *>(s->out_d);
    CK(cudaMemsetAsync(d, 0, sizeof(unsigned long long), s->stream));
    division_selftest_kernel<<<s->num_sms * 8, 256, 0, s->stream>>>(n_pairs, seed, d);
    s->launches++;
    CK(cudaGetLastError());
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, d, sizeof(h), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    *mismatches = (long long)h;
    return BUNMPC_OK;
}

int bunmpc_measure_fp64_peak(bunmpc_solver *s, double *tflops)
{
    if (!s || !tflops) return fail(BUNMPC_ERR_ARG, "peak: null argument");
    CK(cudaSetDevice(s->device));
    const int iters = 4096, threads = 256, blocks = s->num_sms * 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(e0, s->stream));
        fp64_peak_kernel<<<blocks, threads, 0, s->stream>>>(s->out_d, iters, 1.0 + rep);
        CK(cudaEventRecord(e1, s->stream));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * (double)iters * (double)threads * (double)blocks;
        const double tf = flops / (ms * 1e-3) * 1e-12;
        if (rep > 0 && tf > best) best = tf;    // rep 0 is the warm-up
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *tflops = best;
    return BUNMPC_OK;
}

}  // extern "C"
