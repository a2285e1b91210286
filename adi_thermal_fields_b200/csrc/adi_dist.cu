// adi_dist.cu -- the z-slab ADI step sequenced INSIDE the library over NCCL (SURVEY.md 8b "adi_dist_init",
// 8e): one process per GPU, each holding the z planes [z0_r, z1_r) of every array.
//
//   adi_dist_unique_id / adi_dist_init     communicator: rank 0 makes the id, the host carries its 128 bytes to the
//                                          other ranks by whatever means it has (MPI, files, torch.distributed ...)
//   adi_cart_slab_sync_mask                after a mask change: the adjacent ranks' mask planes (neighbour code)
//   adi_cart_slab_step                     one theta-step of adi_step_gpu_coeff (adi3d_gpu_coeff.py:213-230)
//
// Per step: (1) the boundary T planes go to the adjacent ranks (ncclSend / ncclRecv on the library's own
// communication stream) for the explicit stage; x and y sweeps are rank-local; (2) the z sweep is the partitioned
// solve of adi_core.h.  Steady stepping uses its solve-first form, cut into line batches: while batch b's
// (y_first, y_last) pairs are all-gathered on the communication stream, batch b+1 is being solved on the compute
// stream; the ghost corrections of a batch follow as soon as its gather has landed.  After a change of mask,
// packs, dt or theta the two-pass form runs (it also yields the matrix part of the interface relations, gathered
// once and cached), and after `spike_after` steps with the same operands the unit-ghost responses are built.
//
// NCCL is bound at run time (dlopen "libnccl.so.2": in a PyTorch process that is the library torch itself
// loaded), so libadi_b200.so has no link-time dependency on it and single-GPU hosts never touch it.
#include <dlfcn.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "adi_ctx.h"

namespace adi {

int cart_ensure_ghost(adi_ctx *ctx, size_t nlines, cudaStream_t st);
int cart_zsolve0_range(adi_ctx *ctx, double *d_T, double *d_dyn, double dt, double theta, double kappa, double Tinf,
                       size_t line0, size_t nlb, cudaStream_t st);
int cart_zapply_range(adi_ctx *ctx, double *d_T, const double *d_dyn_all, const double *d_stat_all, const double *d_vC,
                      const double *d_wC, const int *d_Kv, const int *d_Kw, int kmax, size_t line0, size_t nlb,
                      cudaStream_t st);
int cart_prof_mark(adi_ctx *ctx, int slot, cudaStream_t st);
int cart_prepare(adi_ctx *ctx, cudaStream_t st);
int cart_explicit_faces(adi_ctx *ctx, const double *d_Tin, double *d_Tout, const double *d_Tlo, const double *d_Thi,
                        double dt, double theta, double kappa, cudaStream_t st);
int cart_step_xy_deferred(adi_ctx *ctx, const double *d_Tin, double *d_Tout, const double *d_Tlo, const double *d_Thi,
                          double dt, double theta, double kappa, double Tinf, cudaStream_t st, cudaEvent_t halo_ready);

// the few NCCL entry points used, by their public C signatures (nccl.h; types reduced to what the ABI needs)
typedef struct ncclComm *nccl_comm_t;
typedef struct { char internal[128]; } nccl_uid_t;
enum { NCCL_UINT8 = 1, NCCL_FLOAT64 = 8 };   // ncclDataType_t values of ncclUint8 / ncclFloat64 (nccl.h)

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_uid_t *) = nullptr;
    int (*CommInitRank)(nccl_comm_t *, int, nccl_uid_t, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

static NcclApi g_nccl;

static int nccl_load()
{
    if (g_nccl.lib) return ADI_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        set_error(std::string("adi_dist: cannot load NCCL (libnccl.so.2): ") + dlerror());
        return ADI_ESTATE;
    }
    NcclApi a;
    a.lib = h;
#define ADI_SYM(field, name)                                                         \
    *(void **)(&a.field) = dlsym(h, name);                                           \
    if (!a.field) { set_error(std::string("adi_dist: NCCL symbol missing: ") + name); return ADI_ESTATE; }
    ADI_SYM(GetUniqueId, "ncclGetUniqueId")
    ADI_SYM(CommInitRank, "ncclCommInitRank")
    ADI_SYM(CommDestroy, "ncclCommDestroy")
    ADI_SYM(AllGather, "ncclAllGather")
    ADI_SYM(Send, "ncclSend")
    ADI_SYM(Recv, "ncclRecv")
    ADI_SYM(GroupStart, "ncclGroupStart")
    ADI_SYM(GroupEnd, "ncclGroupEnd")
    ADI_SYM(GetErrorString, "ncclGetErrorString")
#undef ADI_SYM
    g_nccl = a;
    return ADI_OK;
}

static int nccl_fail(int r, const char *what)
{
    set_error(std::string(what) + ": NCCL error " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    return ADI_ECUDA;
}
#define ADI_NCCL(call)                                        \
    do {                                                      \
        int _r = (call);                                      \
        if (_r != 0) return adi::nccl_fail(_r, #call);        \
    } while (0)

constexpr int MAX_BATCH = 16;

struct DistState {
    nccl_comm_t comm = nullptr;
    bool own_comm = false;
    int rank = 0, nranks = 1;
    cudaStream_t cs = nullptr;                   // communication stream
    cudaEvent_t ev_c = nullptr, ev_x = nullptr;  // compute -> comm, comm -> compute
    cudaEvent_t ev_solved[MAX_BATCH] = {}, ev_gathered[MAX_BATCH] = {};
    // exchange buffers of the bound grid
    size_t nl = 0;
    uint8_t *m_send[2] = {nullptr, nullptr}, *m_recv[2] = {nullptr, nullptr};
    double *t_send[2] = {nullptr, nullptr}, *t_recv[2] = {nullptr, nullptr};
    double *dyn = nullptr, *stat = nullptr, *dyn_all = nullptr, *stat_all = nullptr;
    // cache of the matrix part of the interface relations, and the unit-ghost responses
    bool stat_valid = false;
    double k_dt = 0, k_theta = 0, k_kappa = 0;
    long k_epoch = -1;
    int uses = 0;
    int spike_state = 0;      // 0 not built, 1 in use, -1 responses reach further than kmax: two-pass form stays
    int kmax = 0;
    double *vC = nullptr, *wC = nullptr;
    int *Kv = nullptr, *Kw = nullptr;
    // options
    int nbatch = 4, spike_after = 2, spike_kmax = 32, overlap_halo = 0;
    long batch_min = 1 << 20; // lines: smaller batches are not worth a collective of their own (measured at 262144 lines, N = 2: one batch 1.977 ms, four 2.02 ms per step)
    double spike_thr = 0x1p-60;   // responses below this fraction of the ghost value are dropped: under 1/100 of an ulp of a field of the ghost's magnitude
    long steps_two_pass = 0, steps_solve_first = 0;
};

static void free_buffers(DistState *d)
{
    for (int i = 0; i < 2; ++i) {
        cudaFree(d->m_send[i]); cudaFree(d->m_recv[i]); cudaFree(d->t_send[i]); cudaFree(d->t_recv[i]);
        d->m_send[i] = d->m_recv[i] = nullptr; d->t_send[i] = d->t_recv[i] = nullptr;
    }
    cudaFree(d->dyn); cudaFree(d->stat); cudaFree(d->dyn_all); cudaFree(d->stat_all);
    cudaFree(d->vC); cudaFree(d->wC); cudaFree(d->Kv); cudaFree(d->Kw);
    d->dyn = d->stat = d->dyn_all = d->stat_all = d->vC = d->wC = nullptr;
    d->Kv = d->Kw = nullptr;
    d->nl = 0; d->kmax = 0;
    d->stat_valid = false; d->spike_state = 0;
}

void dist_release(adi_ctx *ctx)
{
    DistState *d = ctx->dist;
    if (!d) return;
    free_buffers(d);
    if (d->comm && d->own_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d->comm);
    if (d->cs) cudaStreamDestroy(d->cs);
    if (d->ev_c) cudaEventDestroy(d->ev_c);
    if (d->ev_x) cudaEventDestroy(d->ev_x);
    for (int i = 0; i < MAX_BATCH; ++i) {
        if (d->ev_solved[i]) cudaEventDestroy(d->ev_solved[i]);
        if (d->ev_gathered[i]) cudaEventDestroy(d->ev_gathered[i]);
    }
    delete d;
    ctx->dist = nullptr;
}

static int ensure_buffers(adi_ctx *ctx)
{
    DistState *d = ctx->dist;
    const size_t nl = (size_t)ctx->nx * ctx->ny;
    if (d->nl == nl && d->dyn) return ADI_OK;
    ADI_CUDA(cudaDeviceSynchronize());
    free_buffers(d);
    const size_t n1 = std::max<size_t>(nl, 1);
    for (int i = 0; i < 2; ++i) {
        ADI_CUDA(cudaMalloc(&d->m_send[i], n1)); ADI_CUDA(cudaMalloc(&d->m_recv[i], n1));
        ADI_CUDA(cudaMalloc(&d->t_send[i], n1 * 8)); ADI_CUDA(cudaMalloc(&d->t_recv[i], n1 * 8));
    }
    ADI_CUDA(cudaMalloc(&d->dyn, 2 * n1 * 8));
    ADI_CUDA(cudaMalloc(&d->stat, 4 * n1 * 8));
    ADI_CUDA(cudaMalloc(&d->dyn_all, (size_t)d->nranks * 2 * n1 * 8));
    ADI_CUDA(cudaMalloc(&d->stat_all, (size_t)d->nranks * 4 * n1 * 8));
    d->nl = nl;
    return ADI_OK;
}

static int check_dist(adi_ctx *ctx, const char *who)
{
    if (!ctx) { set_error(std::string(who) + ": ctx is NULL"); return ADI_EINVAL; }
    if (!ctx->dist || !ctx->dist->comm) { set_error(std::string(who) + ": adi_dist_init has not been called"); return ADI_ESTATE; }
    if (!ctx->cart_bound) { set_error(std::string(who) + ": adi_cart_bind has not been called"); return ADI_ESTATE; }
    ADI_CUDA(cudaSetDevice(ctx->device));
    return ADI_OK;
}

// planes to / from the adjacent ranks: lo plane down, hi plane up (on the communication stream)
template <typename T>
static int exchange_planes(DistState *d, T *const send[2], T *const recv[2], size_t count, int dtype)
{
    ADI_NCCL(g_nccl.GroupStart());
    if (d->rank + 1 < d->nranks) {
        ADI_NCCL(g_nccl.Send(send[1], count, dtype, d->rank + 1, d->comm, d->cs));
        ADI_NCCL(g_nccl.Recv(recv[1], count, dtype, d->rank + 1, d->comm, d->cs));
    }
    if (d->rank > 0) {
        ADI_NCCL(g_nccl.Send(send[0], count, dtype, d->rank - 1, d->comm, d->cs));
        ADI_NCCL(g_nccl.Recv(recv[0], count, dtype, d->rank - 1, d->comm, d->cs));
    }
    ADI_NCCL(g_nccl.GroupEnd());
    return ADI_OK;
}

}  // namespace adi

using namespace adi;

extern "C" {

int adi_dist_unique_id(void *id128)
{
    if (!id128) { set_error("adi_dist_unique_id: NULL argument"); return ADI_EINVAL; }
    int rc = nccl_load();
    if (rc) return rc;
    nccl_uid_t id;
    ADI_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return ADI_OK;
}

static int dist_setup(adi_ctx *ctx, nccl_comm_t comm, bool own, int rank, int nranks)
{
    DistState *d = new DistState();
    d->comm = comm; d->own_comm = own; d->rank = rank; d->nranks = nranks;
    ctx->dist = d;
    ADI_CUDA(cudaStreamCreateWithFlags(&d->cs, cudaStreamNonBlocking));
    ADI_CUDA(cudaEventCreateWithFlags(&d->ev_c, cudaEventDisableTiming));
    ADI_CUDA(cudaEventCreateWithFlags(&d->ev_x, cudaEventDisableTiming));
    for (int i = 0; i < MAX_BATCH; ++i) {
        ADI_CUDA(cudaEventCreateWithFlags(&d->ev_solved[i], cudaEventDisableTiming));
        ADI_CUDA(cudaEventCreateWithFlags(&d->ev_gathered[i], cudaEventDisableTiming));
    }
    return ADI_OK;
}

int adi_dist_init(adi_ctx *ctx, const void *id128, int rank, int nranks)
{
    if (!ctx || !id128 || nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks) {
        set_error("adi_dist_init: need a context, the 128-byte id of adi_dist_unique_id and 0 <= rank < nranks <= 16");
        return ADI_EINVAL;
    }
    int rc = nccl_load();
    if (rc) return rc;
    ADI_CUDA(cudaSetDevice(ctx->device));
    dist_release(ctx);
    nccl_uid_t id;
    memcpy(&id, id128, sizeof(id));
    nccl_comm_t comm = nullptr;
    ADI_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
    return dist_setup(ctx, comm, true, rank, nranks);
}

int adi_dist_init_comm(adi_ctx *ctx, void *nccl_comm, int rank, int nranks)
{
    if (!ctx || !nccl_comm || nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks) {
        set_error("adi_dist_init_comm: need a context, an ncclComm_t and 0 <= rank < nranks <= 16");
        return ADI_EINVAL;
    }
    int rc = nccl_load();
    if (rc) return rc;
    ADI_CUDA(cudaSetDevice(ctx->device));
    dist_release(ctx);
    return dist_setup(ctx, (nccl_comm_t)nccl_comm, false, rank, nranks);
}

int adi_dist_comm(adi_ctx *ctx, void **nccl_comm)
{
    if (!ctx || !ctx->dist || !nccl_comm) { set_error("adi_dist_comm: adi_dist_init has not been called"); return ADI_ESTATE; }
    *nccl_comm = (void *)ctx->dist->comm;
    return ADI_OK;
}

int adi_dist_destroy(adi_ctx *ctx)
{
    if (!ctx) return ADI_EINVAL;
    ADI_CUDA(cudaSetDevice(ctx->device));
    ADI_CUDA(cudaDeviceSynchronize());
    dist_release(ctx);
    return ADI_OK;
}

int adi_dist_info(adi_ctx *ctx, int *rank, int *nranks, long *steps_two_pass, long *steps_solve_first)
{
    if (!ctx || !ctx->dist) { set_error("adi_dist_info: adi_dist_init has not been called"); return ADI_ESTATE; }
    if (rank) *rank = ctx->dist->rank;
    if (nranks) *nranks = ctx->dist->nranks;
    if (steps_two_pass) *steps_two_pass = ctx->dist->steps_two_pass;
    if (steps_solve_first) *steps_solve_first = ctx->dist->steps_solve_first;
    return ADI_OK;
}

int adi_dist_set_option(adi_ctx *ctx, const char *name, long value)
{
    if (!ctx || !ctx->dist || !name) { set_error("adi_dist_set_option: adi_dist_init has not been called"); return ADI_ESTATE; }
    DistState *d = ctx->dist;
    if (!strcmp(name, "batches")) d->nbatch = (int)std::min<long>(std::max<long>(value, 1), MAX_BATCH);
    else if (!strcmp(name, "batch_min_lines")) d->batch_min = std::max<long>(value, 32);
    else if (!strcmp(name, "overlap_halo")) d->overlap_halo = value ? 1 : 0;   // 1: the explicit stage runs while the T planes travel, the face cells follow them on the communication stream (default 0: measured slower at N = 2, r02t / r02u)
    else if (!strcmp(name, "spike_thr_log2")) { d->spike_thr = std::ldexp(1.0, (int)std::min<long>(std::max<long>(value, -200), -30)); d->spike_state = 0; d->stat_valid = false; }
    else if (!strcmp(name, "spike_after")) d->spike_after = (int)value;       // < 0: never leave the two-pass form
    else if (!strcmp(name, "spike_kmax")) { d->spike_kmax = (int)std::max<long>(value, 1); d->spike_state = 0; }
    else { set_error(std::string("adi_dist_set_option: unknown option ") + name); return ADI_EINVAL; }
    return ADI_OK;
}

int adi_cart_slab_sync_mask(adi_ctx *ctx, void *stream)
{
    int rc = check_dist(ctx, "adi_cart_slab_sync_mask");
    if (rc) return rc;
    if (!ctx->d_mask) { set_error("adi_cart_slab_sync_mask: no mask bound"); return ADI_ESTATE; }
    DistState *d = ctx->dist;
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = adi_cart_set_slab(ctx, d->rank, d->nranks))) return rc;
    if ((rc = ensure_buffers(ctx))) return rc;
    const size_t nl = d->nl;
    if (nl && ctx->nz) {
        if ((rc = adi_cart_pack_zplanes(ctx, ctx->d_mask, 1, d->m_send[0], d->m_send[1], stream))) return rc;
        ADI_CUDA(cudaEventRecord(d->ev_c, st));
        ADI_CUDA(cudaStreamWaitEvent(d->cs, d->ev_c, 0));
        if ((rc = exchange_planes<uint8_t>(d, d->m_send, d->m_recv, nl, NCCL_UINT8))) return rc;
        ADI_CUDA(cudaEventRecord(d->ev_x, d->cs));
        ADI_CUDA(cudaStreamWaitEvent(st, d->ev_x, 0));
    }
    return adi_cart_set_mask_halo(ctx, d->rank > 0 ? d->m_recv[0] : nullptr, d->rank + 1 < d->nranks ? d->m_recv[1] : nullptr);
}

int adi_cart_slab_step(adi_ctx *ctx, const double *d_Tin, double *d_Tout, double dt, double theta, double kappa,
                       double Tinf, void *stream)
{
    int rc = check_dist(ctx, "adi_cart_slab_step");
    if (rc) return rc;
    if (!d_Tin || !d_Tout || d_Tin == d_Tout) {
        set_error("adi_cart_slab_step: Tin/Tout must be distinct device arrays");
        return ADI_EINVAL;
    }
    DistState *d = ctx->dist;
    cudaStream_t st = (cudaStream_t)stream;
    if (d->nranks == 1) return adi_cart_step(ctx, d_Tin, d_Tout, dt, theta, kappa, Tinf, stream);
    if (ctx->slab_nranks != d->nranks || ctx->slab_rank != d->rank) {
        set_error("adi_cart_slab_step: adi_cart_slab_sync_mask has not been called for this grid");
        return ADI_ESTATE;
    }
    if ((rc = ensure_buffers(ctx))) return rc;
    const size_t nl = d->nl;
    const int nz = ctx->nz, R = d->nranks;
    if (!nl || !nz) return ADI_OK;
    const bool lo_ok = d->rank > 0, hi_ok = d->rank + 1 < R;

    // (1) T planes for the explicit stage (beta = 0 at theta = 1: no explicit stage, no halo).  The planes travel on
    // the communication stream while the explicit stage runs; only the cells on the slab faces wait for them.
    cudaEvent_t halo_ready = nullptr;
    if (theta != 1.0) {
        const bool overlap = d->overlap_halo && !ctx->opt_fuse;
        if (overlap && (rc = cart_prepare(ctx, st))) return rc;      // the neighbour code must exist before ev_c
        if ((rc = adi_cart_pack_zplanes(ctx, d_Tin, 8, d->t_send[0], d->t_send[1], stream))) return rc;
        ADI_CUDA(cudaEventRecord(d->ev_c, st));
        ADI_CUDA(cudaStreamWaitEvent(d->cs, d->ev_c, 0));
        if ((rc = exchange_planes<double>(d, d->t_send, d->t_recv, nl, NCCL_FLOAT64))) return rc;
        if (overlap) {
            // the face cells follow the planes on the communication stream, beside the explicit stage of everything else
            if ((rc = cart_explicit_faces(ctx, d_Tin, d_Tout, lo_ok ? d->t_recv[0] : nullptr, hi_ok ? d->t_recv[1] : nullptr, dt,
                                          theta, kappa, d->cs)))
                return rc;
            ADI_CUDA(cudaEventRecord(d->ev_x, d->cs));
            halo_ready = d->ev_x;
        } else {
            ADI_CUDA(cudaEventRecord(d->ev_x, d->cs));
            ADI_CUDA(cudaStreamWaitEvent(st, d->ev_x, 0));
        }
    }
    if ((rc = cart_step_xy_deferred(ctx, d_Tin, d_Tout, lo_ok ? d->t_recv[0] : nullptr, hi_ok ? d->t_recv[1] : nullptr, dt,
                                    theta, kappa, Tinf, st, halo_ready)))
        return rc;

    // (2) z sweep.  The matrix part of the interface relations only changes with mask, packs, dt or theta.
    const bool fresh = !(d->stat_valid && d->k_dt == dt && d->k_theta == theta && d->k_kappa == kappa &&
                         d->k_epoch == ctx->operand_epoch);
    if (fresh) {
        d->spike_state = 0;
        d->uses = 0;
    } else {
        d->uses++;
        if (d->spike_state == 0 && d->spike_after >= 0 && d->uses >= d->spike_after) {
            // a purely local decision: both forms hand the same relation to the other ranks
            const int kmax = std::min(d->spike_kmax, nz);
            if (d->kmax != kmax || !d->vC) {
                ADI_CUDA(cudaStreamSynchronize(st));
                cudaFree(d->vC); cudaFree(d->wC); cudaFree(d->Kv); cudaFree(d->Kw);
                d->vC = d->wC = nullptr; d->Kv = d->Kw = nullptr;
                ADI_CUDA(cudaMalloc(&d->vC, nl * kmax * 8)); ADI_CUDA(cudaMalloc(&d->wC, nl * kmax * 8));
                ADI_CUDA(cudaMalloc(&d->Kv, nl * sizeof(int))); ADI_CUDA(cudaMalloc(&d->Kw, nl * sizeof(int)));
                d->kmax = kmax;
            }
            double *scratch = nullptr;
            ADI_CUDA(cudaMalloc(&scratch, nl * (size_t)nz * 8));
            int mk0 = 0, mk1 = 0;
            rc = adi_cart_zsweep_spike(ctx, scratch, 0, kmax, d->spike_thr, d->vC, d->Kv, &mk0, dt, theta, kappa, stream);
            if (!rc) rc = adi_cart_zsweep_spike(ctx, scratch, 1, kmax, d->spike_thr, d->wC, d->Kw, &mk1, dt, theta, kappa, stream);
            cudaFree(scratch);
            if (rc) return rc;
            d->spike_state = (mk0 > kmax || mk1 > kmax) ? -1 : 1;
        }
    }
    if (d->spike_state == 1) {
        // solve first, correct at the faces; line batches overlap the all-gather of one with the solve of the next
        if ((rc = cart_ensure_ghost(ctx, nl, st))) return rc;
        int nb = std::max(1, std::min(d->nbatch, MAX_BATCH));
        size_t per = ((nl + nb - 1) / nb + 31) & ~(size_t)31;
        if (per < (size_t)d->batch_min) per = std::min<size_t>(nl, (size_t)d->batch_min);    // tiny grids: not worth splitting
        nb = (int)((nl + per - 1) / per);
        for (int b = 0; b < nb; ++b) {
            const size_t l0 = (size_t)b * per, n = std::min(per, nl - l0);
            if ((rc = cart_zsolve0_range(ctx, d_Tout, d->dyn + 2 * l0, dt, theta, kappa, Tinf, l0, n, st))) return rc;
            ADI_CUDA(cudaEventRecord(d->ev_solved[b], st));
            ADI_CUDA(cudaStreamWaitEvent(d->cs, d->ev_solved[b], 0));
            ADI_NCCL(g_nccl.AllGather(d->dyn + 2 * l0, d->dyn_all + (size_t)R * 2 * l0, 2 * n, NCCL_FLOAT64, d->comm, d->cs));
            ADI_CUDA(cudaEventRecord(d->ev_gathered[b], d->cs));
        }
        for (int b = 0; b < nb; ++b) {
            const size_t l0 = (size_t)b * per, n = std::min(per, nl - l0);
            ADI_CUDA(cudaStreamWaitEvent(st, d->ev_gathered[b], 0));
            if ((rc = cart_zapply_range(ctx, d_Tout, d->dyn_all + (size_t)R * 2 * l0, d->stat_all, d->vC, d->wC, d->Kv, d->Kw,
                                        d->kmax, l0, n, st)))
                return rc;
        }
        d->steps_solve_first++;
        return cart_prof_mark(ctx, 4, st);
    }
    // two-pass form
    if ((rc = adi_cart_zsweep_reduce(ctx, d_Tout, d->dyn, fresh ? d->stat : nullptr, dt, theta, kappa, Tinf, stream))) return rc;
    ADI_CUDA(cudaEventRecord(d->ev_c, st));
    ADI_CUDA(cudaStreamWaitEvent(d->cs, d->ev_c, 0));
    ADI_NCCL(g_nccl.AllGather(d->dyn, d->dyn_all, 2 * nl, NCCL_FLOAT64, d->comm, d->cs));
    if (fresh) ADI_NCCL(g_nccl.AllGather(d->stat, d->stat_all, 4 * nl, NCCL_FLOAT64, d->comm, d->cs));
    ADI_CUDA(cudaEventRecord(d->ev_x, d->cs));
    ADI_CUDA(cudaStreamWaitEvent(st, d->ev_x, 0));
    if (fresh) {
        d->stat_valid = true;
        d->k_dt = dt; d->k_theta = theta; d->k_kappa = kappa; d->k_epoch = ctx->operand_epoch;
    }
    d->steps_two_pass++;
    return adi_cart_zsweep_finish(ctx, d_Tout, d->dyn_all, d->stat_all, dt, theta, kappa, Tinf, stream);
}

}  // extern "C"
