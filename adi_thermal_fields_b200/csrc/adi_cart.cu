// adi_cart.cu -- Cartesian entry points of the C ABI: operand binding, variant
// selection and launch of the kernels in adi_cart.cuh.
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#define ADI_CART_MISC_KERNELS
#include "adi_cart.cuh"
#include "adi_ctx.h"

using namespace adi;

namespace {


int check_cart(adi_ctx *ctx, const char *who)
{
    if (!ctx) {
        set_error(std::string(who) + ": ctx is NULL");
        return ADI_EINVAL;
    }
    if (!ctx->cart_bound) {
        set_error(std::string(who) + ": adi_cart_bind has not been called");
        return ADI_ESTATE;
    }
    ADI_CUDA(cudaSetDevice(ctx->device));   // the context's device, whatever the caller made current
    return ADI_OK;
}

int ensure_code(adi_ctx *ctx, cudaStream_t st)
{
    if (!ctx->code_dirty) return ADI_OK;
    if (!ctx->d_mask) {
        set_error("adi_cart_step: no mask bound (adi_cart_set_mask)");
        return ADI_ESTATE;
    }
    const size_t n = (size_t)ctx->nx * ctx->ny * ctx->nz;
    // one code array when the three packs share their Dirichlet mask, else one per axis
    const bool shared = ctx->pack[0].dirm == ctx->pack[1].dirm && ctx->pack[1].dirm == ctx->pack[2].dirm;
    const int ncodes = shared ? 1 : 3;
    if (ctx->code_cells != n) {
        for (int a = 0; a < 3; ++a)
            if (ctx->code_buf[a]) { cudaFree(ctx->code_buf[a]); ctx->code_buf[a] = nullptr; }
        ctx->code_cells = n;
    }
    for (int a = 0; a < ncodes; ++a)
        if (!ctx->code_buf[a]) ADI_CUDA(cudaMalloc(&ctx->code_buf[a], n ? n : 1));
    if (n) {
        const int threads = 256;
        const int blocks = (int)std::min<size_t>((n + threads - 1) / threads, 148 * 64);
        ctx->maskv_used &= ~3;
        ctx->ztop = -1; ctx->ztop_pending = false;     // unknown unless the word form reports it
        if (!ctx->d_ztop) {
            ADI_CUDA(cudaMalloc(&ctx->d_ztop, 3 * sizeof(int)));
            ADI_CUDA(cudaMallocHost(&ctx->h_ztop, 3 * sizeof(int)));
        }
        for (int a = 0; a < ncodes; ++a) {
            // word form: 16 cells per thread (adi_mask_core.h) when every z line starts on a 16-byte boundary
            const bool wordform = ctx->opt_maskv && ctx->nz % 16 == 0 &&
                                  (((uintptr_t)ctx->d_mask | (uintptr_t)ctx->pack[a].dirm | (uintptr_t)ctx->code_buf[a]) & 15) == 0;
            if (wordform) {
                const int vblocks = (int)std::min<size_t>((n / 16 + threads - 1) / threads, 148 * 32);
                if (a == 0) ADI_CUDA(cudaMemsetAsync(ctx->d_ztop, 0, 3 * sizeof(int), st));
                k_build_code_v<<<vblocks, threads, 0, st>>>(ctx->d_mask, ctx->pack[a].dirm, ctx->code_buf[a],
                                                            ctx->nx, ctx->ny, ctx->nz, ctx->d_mask_lo, ctx->d_mask_hi,
                                                            a == 0 ? ctx->d_ztop : nullptr);
                if (a == 0) {   // top of the part, read when the z sweep first needs it (launch_sweep_zt)
                    ADI_CUDA(cudaMemcpyAsync(ctx->h_ztop, ctx->d_ztop, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
                    ctx->ztop_pending = true;
                }
                ctx->maskv_used |= 1;
            } else {
                k_build_code<<<blocks, threads, 0, st>>>(ctx->d_mask, ctx->pack[a].dirm, ctx->code_buf[a],
                                                         ctx->nx, ctx->ny, ctx->nz, ctx->d_mask_lo, ctx->d_mask_hi);
            }
            ctx->launches++;
        }
        ADI_CUDA(cudaGetLastError());
    }
    for (int a = 0; a < 3; ++a) ctx->code[a] = ctx->code_buf[shared ? 0 : a];
    // transposed copies for the x / y sweeps: line axis fastest, lines padded to a multiple of 32
    for (int a = 0; a < 2; ++a) {
        const int len = a == 0 ? ctx->nx : ctx->ny, batch = a == 0 ? ctx->ny : ctx->nx;
        const int npad = (len + 31) / 32 * 32;
        const size_t bytes = (size_t)batch * ctx->nz * npad;
        if (!ctx->opt_xy2 || !n || len > 2048) { ctx->npadT[a] = 0; continue; }
        if (ctx->codeT_bytes[a] != bytes || ctx->npadT[a] != npad) {
            if (ctx->codeT[a]) { ADI_CUDA(cudaFree(ctx->codeT[a])); ctx->codeT[a] = nullptr; }
            ADI_CUDA(cudaMalloc(&ctx->codeT[a], bytes));
            ADI_CUDA(cudaMemsetAsync(ctx->codeT[a], 0, bytes, st));
            ctx->codeT_bytes[a] = bytes; ctx->npadT[a] = npad;
        }
        const size_t snx = (size_t)ctx->ny * ctx->nz;
        const size_t sb = a == 0 ? (size_t)ctx->nz : snx, sr = a == 0 ? snx : (size_t)ctx->nz;
        // word form: 128 x 128 byte tiles (adi_mask_core.h) when rows are word aligned
        if (ctx->opt_maskv && ctx->nz % 4 == 0 && (((uintptr_t)ctx->code[a] | (uintptr_t)ctx->codeT[a]) & 3) == 0) {
            TrArgs t;
            t.src = ctx->code[a]; t.dst = ctx->codeT[a]; t.n = len; t.nz = ctx->nz; t.npad = npad; t.sb = sb; t.sr = sr;
            dim3 vgrid((unsigned)((ctx->nz + 127) / 128), (unsigned)((len + 127) / 128), (unsigned)std::min(batch, 65535));
            k_transpose_code_v<<<vgrid, 256, 0, st>>>(t, batch);
            ctx->maskv_used |= 2;
        } else {
            dim3 tgrid((unsigned)((ctx->nz + 31) / 32), (unsigned)((len + 31) / 32), (unsigned)std::min(batch, 65535));
            k_transpose_code<<<tgrid, dim3(32, 8), 0, st>>>(ctx->code[a], ctx->codeT[a], len, ctx->nz, npad, batch, sb, sr);
        }
        ctx->launches++;
        ADI_CUDA(cudaGetLastError());
    }
    ctx->code_dirty = false;
    ctx->sparse_dirty = true;
    for (int a = 0; a < 3; ++a) ctx->tiles[a].valid = false;
    return ADI_OK;
}

// After a pack or mask change: which dense coefficient fields vanish away from the surface along their
// own axis?  (precompute_coeff_packs_unified output always does, adi3d_numba_coeff.py:93-99.)  One pass over
// the fields and one small read-back; the sweeps then skip the coefficient reads of interior cells.
int ensure_sparse(adi_ctx *ctx, cudaStream_t st)
{
    if (!ctx->sparse_dirty) return ADI_OK;
    for (int a = 0; a < 3; ++a) ctx->sparse[a] = false;
    ctx->sparse_dirty = false;
    const size_t n = (size_t)ctx->nx * ctx->ny * ctx->nz;
    if (!ctx->opt_sparse || ctx->scalar_robin || n == 0) return ADI_OK;
    if (!ctx->pack[0].coeff && !ctx->pack[1].coeff && !ctx->pack[2].coeff) return ADI_OK;
    if (!ctx->d_viol) {
        ADI_CUDA(cudaMalloc(&ctx->d_viol, 3 * sizeof(unsigned long long)));
        ADI_CUDA(cudaMallocHost(&ctx->h_viol, 3 * sizeof(unsigned long long)));
    }
    SparseCheckArgs c;
    bool work = false;
    for (int a = 0; a < 3; ++a) {
        // x, y and (second-generation z sweep, option zt) z; the first-generation z sweep gives a line to one warp,
        // where fetching the coefficients of exposed cells costs a second memory round trip (0.63 -> 0.76 ms at 512^3)
        const bool want = a < 2 || ctx->opt_zt;
        const bool trusted = (ctx->sparse_trust >> a) & 1;
        c.coeff[a] = (want && !trusted) ? ctx->pack[a].coeff : nullptr;
        c.code[a] = ctx->code[a];
        if (want && trusted && ctx->pack[a].coeff) ctx->sparse[a] = true;
        work = work || c.coeff[a] != nullptr;
    }
    if (!work) return ADI_OK;
    c.viol = ctx->d_viol;
    ADI_CUDA(cudaMemsetAsync(ctx->d_viol, 0, 3 * sizeof(unsigned long long), st));
    const int threads = 256;
    const int blocks = (int)std::min<size_t>((n + threads - 1) / threads, 148 * 16);
    k_check_sparse<<<blocks, threads, 0, st>>>(c, n);
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    ADI_CUDA(cudaMemcpyAsync(ctx->h_viol, ctx->d_viol, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    ADI_CUDA(cudaStreamSynchronize(st));
    for (int a = 0; a < 3; ++a)
        if (c.coeff[a]) ctx->sparse[a] = ctx->h_viol[a] == 0ull;
    return ADI_OK;
}

}  // namespace

// The extent of the part as the last code build reduced it (k_build_code_v): ctx->ztop (z + 1 of the highest active
// cell; -1 unknown), ctx->xlo / ctx->xhi (first x plane with an active cell / last + 1).  Waits for the small copy the
// first time it is asked after a mask change.
int adi::part_extent(adi_ctx *ctx, cudaStream_t st)
{
    if (ctx->ztop_pending) {
        ADI_CUDA(cudaStreamSynchronize(st));
        ctx->ztop = ctx->h_ztop[0];
        ctx->xlo = ctx->h_ztop[0] > 0 ? ctx->nx - ctx->h_ztop[1] : 0;
        ctx->xhi = ctx->h_ztop[0] > 0 ? ctx->h_ztop[2] : 0;
        ctx->ztop_pending = false;
    }
    return ADI_OK;
}

// Active-tile list of a sweep axis for tiles of KT rows (see k_tile_flags): flags on the device, compaction on the
// host (one small read-back per mask change), ascending order so that neighbouring blocks keep touching
// neighbouring memory.
int adi::ensure_tiles(adi_ctx *ctx, int axis, int KT, cudaStream_t st, const int **list, int *nactive, int *tiles_nx)
{
    *list = nullptr; *nactive = 0; *tiles_nx = 0;
    if (!ctx->opt_tiles || KT < 1) return ADI_OK;
    const uint8_t *base;
    size_t unit;
    int inner, outer;
    if (axis == 2) {
        base = ctx->code[2]; unit = (size_t)ctx->nz; inner = (int)std::min<size_t>((size_t)ctx->nx * ctx->ny, 0x7fffffff); outer = 1;
        if ((size_t)ctx->nx * ctx->ny > 0x7fffffffull) return ADI_OK;
    } else {
        if (!ctx->codeT[axis] || ctx->npadT[axis] <= 0) return ADI_OK;
        base = ctx->codeT[axis]; unit = (size_t)ctx->npadT[axis]; inner = ctx->nz; outer = axis == 0 ? ctx->ny : ctx->nx;
    }
    const int nti = (inner + KT - 1) / KT;
    const long long total = (long long)nti * outer;
    if (total < 1 || total > 0x7fffffffll) return ADI_OK;
    TileList &L = ctx->tiles[axis];
    if (!(L.valid && L.kt == KT && L.total == (int)total)) {
        const size_t n = (size_t)total;
        if (ctx->tflags_cap < n) {
            if (ctx->d_tflags) { ADI_CUDA(cudaFree(ctx->d_tflags)); ADI_CUDA(cudaFreeHost(ctx->h_tflags)); }
            ctx->d_tflags = nullptr; ctx->h_tflags = nullptr;
            ADI_CUDA(cudaMalloc(&ctx->d_tflags, n));
            ADI_CUDA(cudaMallocHost(&ctx->h_tflags, n));
            ctx->tflags_cap = n;
        }
        const unsigned ulo = axis == 0 ? CB_XM : CB_YM, uhi = axis == 0 ? CB_XP : (axis == 1 ? CB_YP : 0u);
        k_tile_flags<<<(unsigned)std::min<size_t>(n, 148 * 64), 128, 0, st>>>(base, unit, KT, inner, (int)total, ctx->d_tflags,
                                                                               axis == 2 ? 0u : ulo, uhi,
                                                                               axis == 0 ? ctx->nx : ctx->ny);
        ctx->launches++;
        ADI_CUDA(cudaGetLastError());
        ADI_CUDA(cudaMemcpyAsync(ctx->h_tflags, ctx->d_tflags, n, cudaMemcpyDeviceToHost, st));
        ADI_CUDA(cudaStreamSynchronize(st));
        std::vector<int> ids, uni, gen;
        ids.reserve(n);
        for (size_t t = 0; t < n; ++t) {
            const uint8_t f = ctx->h_tflags[t];
            if (f & 1) {
                ids.push_back((int)t);
                ((f & 2) ? uni : gen).push_back((int)t);
            }
        }
        if (L.cap < std::max<size_t>(ids.size(), 1)) {
            if (L.d) { ADI_CUDA(cudaFree(L.d)); ADI_CUDA(cudaFree(L.d_uni)); ADI_CUDA(cudaFree(L.d_gen)); }
            L.d = L.d_uni = L.d_gen = nullptr;
            L.cap = std::max<size_t>(n, 1);
            ADI_CUDA(cudaMalloc(&L.d, L.cap * sizeof(int)));
            ADI_CUDA(cudaMalloc(&L.d_uni, L.cap * sizeof(int)));
            ADI_CUDA(cudaMalloc(&L.d_gen, L.cap * sizeof(int)));
        }
        if (!ids.empty()) ADI_CUDA(cudaMemcpyAsync(L.d, ids.data(), ids.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        if (!uni.empty()) ADI_CUDA(cudaMemcpyAsync(L.d_uni, uni.data(), uni.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        if (!gen.empty()) ADI_CUDA(cudaMemcpyAsync(L.d_gen, gen.data(), gen.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        ADI_CUDA(cudaStreamSynchronize(st));   // the vectors leave scope
        L.n = (int)ids.size(); L.n_uni = (int)uni.size(); L.n_gen = (int)gen.size();
        L.total = (int)total; L.kt = KT; L.valid = true;
    }
    *tiles_nx = nti;
    if (L.n < L.total) { *list = L.d; *nactive = L.n; }
    return ADI_OK;
}

int adi::ensure_tiles_split(adi_ctx *ctx, int axis, int KT, cudaStream_t st, const int **uni, int *nuni, const int **gen,
                            int *ngen, int *tiles_nx)
{
    *uni = *gen = nullptr; *nuni = *ngen = 0;
    const int *list = nullptr;
    int nact = 0;
    int rc = ensure_tiles(ctx, axis, KT, st, &list, &nact, tiles_nx);
    if (rc) return rc;
    const TileList &L = ctx->tiles[axis];
    if (axis > 1 || !L.valid || L.kt != KT) return ADI_OK;     // (option tiles off, or the grid is too large for lists)
    *uni = L.d_uni; *nuni = L.n_uni; *gen = L.d_gen; *ngen = L.n_gen;
    return ADI_OK;
}

// adi_sweep_x.cu / adi_sweep_y.cu / adi_sweep_z.cu (one translation unit per sweep so that the
// template instantiations compile in parallel)
namespace adi {
int launch_sweep_x(adi_ctx *ctx, const SweepArgs &a, bool dense, bool extra, bool expl, cudaStream_t st);
int launch_sweep_y(adi_ctx *ctx, const SweepArgs &a, bool dense, bool extra, cudaStream_t st);
int launch_sweep_z(adi_ctx *ctx, const SweepArgs &a, bool dense, bool extra, int zmode, cudaStream_t st);
}

namespace {

int ensure_stage(adi_ctx *ctx, size_t cells)
{
    if (ctx->stage_cells >= cells && ctx->stage[0]) return ADI_OK;
    for (int a = 0; a < 2; ++a) {
        if (ctx->stage[a]) cudaFree(ctx->stage[a]);
        ctx->stage[a] = nullptr;
        ADI_CUDA(cudaMalloc(&ctx->stage[a], std::max<size_t>(cells, 1) * sizeof(double)));
    }
    ctx->stage_cells = cells;
    return ADI_OK;
}

}  // namespace

extern "C" {

int adi_cart_bind(adi_ctx *ctx, int nx, int ny, int nz, double dx)
{
    if (!ctx) return ADI_EINVAL;
    if (nx < 0 || ny < 0 || nz < 0 || !(dx > 0.0)) {
        set_error("adi_cart_bind: bad grid");
        return ADI_EINVAL;
    }
    ADI_CUDA(cudaSetDevice(ctx->device));
    ctx->nx = nx; ctx->ny = ny; ctx->nz = nz; ctx->dx = dx;
    ctx->cart_bound = true;
    ctx->d_mask = nullptr;
    ctx->d_mask_lo = ctx->d_mask_hi = nullptr;
    ctx->slab_rank = 0; ctx->slab_nranks = 1;
    for (int a = 0; a < 3; ++a) ctx->pack[a] = Pack();
    ctx->scalar_robin = false;
    ctx->code_dirty = true;
    ctx->sparse_trust = 0;
    return ADI_OK;
}

int adi_cart_set_mask(adi_ctx *ctx, const uint8_t *d_mask)
{
    int rc = check_cart(ctx, "adi_cart_set_mask");
    if (rc) return rc;
    ctx->d_mask = d_mask;
    ctx->code_dirty = true;
    ctx->sparse_trust = 0;
    ctx->operand_epoch++;
    return ADI_OK;
}

int adi_cart_set_pack(adi_ctx *ctx, int axis, const double *d_coeff, const uint8_t *d_dir_mask,
                      const double *d_dir_val, const double *d_qflux)
{
    int rc = check_cart(ctx, "adi_cart_set_pack");
    if (rc) return rc;
    if (axis < 0 || axis > 2) {
        set_error("adi_cart_set_pack: axis must be 0, 1 or 2");
        return ADI_EINVAL;
    }
    if (d_dir_mask && !d_dir_val) {
        set_error("adi_cart_set_pack: dir_mask given without dir_val");
        return ADI_EINVAL;
    }
    Pack &p = ctx->pack[axis];
    if (p.dirm || d_dir_mask) ctx->code_dirty = true;  // Dirichlet bit lives in the code
    p.coeff = d_coeff; p.dirm = d_dir_mask; p.dirv = d_dir_mask ? d_dir_val : nullptr; p.q = d_qflux;
    ctx->scalar_robin = false;
    ctx->sparse_dirty = true;
    ctx->sparse_trust = 0;
    ctx->operand_epoch++;
    return ADI_OK;
}

int adi_cart_set_robin_scalar(adi_ctx *ctx, const double face_coeff[6])
{
    int rc = check_cart(ctx, "adi_cart_set_robin_scalar");
    if (rc) return rc;
    if (!face_coeff) return ADI_EINVAL;
    for (int f = 0; f < 6; ++f) ctx->face_coeff[f] = face_coeff[f];
    for (int a = 0; a < 3; ++a) ctx->pack[a].coeff = nullptr;
    ctx->scalar_robin = true;
    ctx->sparse_dirty = true;
    ctx->sparse_trust = 0;
    ctx->operand_epoch++;
    return ADI_OK;
}

// Sweeps `first`..`last` (0 = explicit stage + x, 1 = y, 2 = z) of one step.
// zmode: 0 whole z lines, 1 z-slab pass 1 (interface relations -> d_iface), 2 z-slab pass 2.
// line0 / nlb: (z sweep only) restrict the sweep to the z lines [line0, line0 + nlb) -- the batches of the
// overlapped multi-GPU z solve; nlb == 0: all lines.  d_iface_* / d_ghost are then the BATCH's arrays ([2][nlb] ...).
static int run_sweeps(adi_ctx *ctx, const double *d_Tin, double *d_Tout, double dt, double theta,
                      double kappa, double Tinf, int first, int last, int zmode, const double *d_Tlo,
                      const double *d_Thi, double *d_iface_dyn, double *d_iface_stat, const double *d_ghost,
                      cudaStream_t st, size_t line0 = 0, size_t nlb = 0, cudaEvent_t halo_ready = nullptr)
{
    const size_t ncell = (size_t)ctx->nx * ctx->ny * ctx->nz;
    if (ncell == 0) return ADI_OK;
    int rc = ensure_code(ctx, st);
    if (rc) return rc;
    if ((rc = ensure_sparse(ctx, st))) return rc;

    // adi3d_numba_coeff.py:291-292,298 -- scalars in the reference's evaluation order
    const double dx = ctx->dx;
    const double gam = kappa * dt / (dx * dx);
    SweepArgs a;
    a.nx = ctx->nx; a.ny = ctx->ny; a.nz = ctx->nz;
    a.k.g = theta * gam;
    a.k.dt = dt;
    a.k.Tinf = Tinf;
    a.k.beta = dt * kappa * (1.0 - theta);
    a.k.invdx2 = 1.0 / (dx * dx);
    a.zlo = d_Tlo; a.zhi = d_Thi; a.iface_dyn = d_iface_dyn; a.iface_stat = d_iface_stat; a.ghost = d_ghost;
    a.codeT = nullptr; a.npad = 0; a.uni = 0; a.tw = 0; a.remap = 0; a.dbg = 0; a.halo_defer = 0;
    a.tiles = nullptr; a.tiles_nx = 0; a.tsplit = 0; a.zpitch = 0; a.code_line = 0; a.zfull = 0; a.line0 = 0;
    a.line_batch = nlb != 0 ? 1 : 0;
    bool expl = a.k.beta != 0.0;
    bool x_in_place = false;
    // halo_ready: d_Tlo / d_Thi are still being received; the explicit stage leaves the cells that need them to
    // k_explicit_faces, which the caller has queued behind the planes on another stream and which records the event
    if (halo_ready && !(expl && first == 0 && !ctx->opt_fuse)) {
        ADI_CUDA(cudaStreamWaitEvent(st, halo_ready, 0));
        halo_ready = nullptr;
    }
    if (expl && first == 0 && !ctx->opt_fuse) {
        // explicit stage as its own streaming pass Tin -> Tout; the x sweep then runs in place
        a.in = d_Tin; a.out = d_Tout; a.code = ctx->code[0];
        a.coeff = nullptr; a.sparse = 0; a.q = nullptr; a.dirv = nullptr;
        const bool vec = (a.nz % 2 == 0) && ((((uintptr_t)d_Tin | (uintptr_t)d_Tout) & 15) == 0) &&
                         (((uintptr_t)a.code & 1) == 0);
        const int VEC = vec ? 2 : 1;
        const int threads = (ctx->opt_eth == 32 || ctx->opt_eth == 64 || ctx->opt_eth == 128) ? (int)ctx->opt_eth : 128;
        const int JT = ctx->opt_ejt > 0 ? (int)std::min<long>(ctx->opt_ejt, 1024) : 16;
        // block order (option eorder): 0 = z chunks fastest, x slowest; 1 = x fastest -- the blocks of neighbouring x
        // planes then run together and find each other's rows (their x-1 / x+1 neighbours) in L2
        const unsigned gz = (unsigned)(((a.nz + VEC - 1) / VEC + threads - 1) / threads), gy = (unsigned)((a.ny + JT - 1) / JT);
        const int xfast = (ctx->opt_eorder && gz <= 65535u) ? 1 : 0;
        dim3 grid = xfast ? dim3((unsigned)a.nx, gy, gz) : dim3(gz, gy, (unsigned)a.nx);
        a.halo_defer = halo_ready ? 1 : 0;
        if (vec) k_explicit<2><<<grid, threads, 0, st>>>(a, JT, xfast);
        else k_explicit<1><<<grid, threads, 0, st>>>(a, JT, xfast);
        ctx->launches++;
        ADI_CUDA(cudaGetLastError());
        a.halo_defer = 0;
        if (halo_ready) ADI_CUDA(cudaStreamWaitEvent(st, halo_ready, 0));   // k_explicit_faces has filled in the face cells
        expl = false;
        x_in_place = true;
    }
    if (first == 0) {
        rc = prof_mark(ctx, 1, st);
        if (rc) return rc;
    }

    for (int axis = first; axis <= last; ++axis) {
        const Pack &p = ctx->pack[axis];
        a.in = (axis == 0 && !x_in_place) ? d_Tin : d_Tout;  // y and z sweeps run in place on Tout
        a.out = d_Tout;
        a.code = ctx->code[axis];
        a.coeff = p.coeff;
        a.sparse = ctx->sparse[axis] ? 1 : 0;
        a.q = zmode == 5 ? nullptr : p.q;        // 5: homogeneous system (unit-ghost response)
        a.dirv = zmode == 5 ? nullptr : p.dirv;
        a.k.h_lo = ctx->scalar_robin ? ctx->face_coeff[2 * axis] : 0.0;
        a.k.h_hi = ctx->scalar_robin ? ctx->face_coeff[2 * axis + 1] : 0.0;
        const bool dense = p.coeff != nullptr;
        const bool extra = p.q != nullptr || p.dirm != nullptr;
        if (axis == 0) {
            // A part under construction occupies the x planes xlo .. xhi-1 only (reduced by the code build): the in-place
            // x sweep runs on that sub-grid -- whole 32-cell chunks, at least 256 planes -- through offset pointers.
            SweepArgs b = a;
            int x0 = 0, x1 = a.nx;
            if (ctx->opt_ztrim && a.in == a.out && a.nx >= 512 && a.nx % 32 == 0 && a.nx == ctx->nx) {
                rc = part_extent(ctx, st);
                if (rc) return rc;
                if (ctx->ztop >= 0) {
                    x0 = ctx->xlo / 32 * 32;
                    x1 = std::min(a.nx, (std::max(ctx->xhi, x0 + 1) + 31) / 32 * 32);
                    if (x1 - x0 < 256) { x0 = std::max(0, std::min(x0, a.nx - 256)); x1 = std::min(a.nx, x0 + 256); }
                }
            }
            if (x1 - x0 < a.nx) {
                const size_t off = (size_t)x0 * a.ny * a.nz;
                b.in += off; b.out += off; b.code += off;
                if (b.coeff) b.coeff += off;
                if (b.q) b.q += off;
                if (b.dirv) b.dirv += off;
                b.nx = x1 - x0; b.line0 = x0;
                ctx->xtrim_used++;
            }
            rc = launch_sweep_x(ctx, b, dense, extra, expl, st);
        }
        else if (axis == 1) rc = launch_sweep_y(ctx, a, dense, extra, st);
        else if (nlb == 0) rc = launch_sweep_z(ctx, a, dense, extra, zmode == 5 ? 2 : zmode, st);
        else {
            SweepArgs b = a;   // the z kernels see a grid of nlb lines
            const size_t off = line0 * (size_t)a.nz;
            b.in += off; b.out += off; b.code += off;
            if (b.coeff) b.coeff += off;
            if (b.q) b.q += off;
            if (b.dirv) b.dirv += off;
            b.nx = 1; b.ny = (int)nlb;
            b.line_batch = 1;
            rc = launch_sweep_z(ctx, b, dense, extra, zmode == 5 ? 2 : zmode, st);
        }
        if (rc) return rc;
        if (zmode == 0 || zmode == 2) {
            rc = prof_mark(ctx, axis + 2, st);
            if (rc) return rc;
        }
    }
    if (ctx->opt_sync_check) ADI_CUDA(cudaStreamSynchronize(st));
    return ADI_OK;
}

int adi_cart_step(adi_ctx *ctx, const double *d_Tin, double *d_Tout, double dt, double theta,
                  double kappa, double Tinf, void *stream)
{
    int rc = check_cart(ctx, "adi_cart_step");
    if (rc) return rc;
    if (!d_Tin || !d_Tout || d_Tin == d_Tout) {
        set_error("adi_cart_step: Tin/Tout must be distinct device arrays");
        return ADI_EINVAL;
    }
    if (ctx->slab_nranks > 1) {
        set_error("adi_cart_step: this context holds a z slab; use adi_cart_step_xy + adi_cart_zsweep_*");
        return ADI_ESTATE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    rc = prof_mark(ctx, 0, st);
    if (rc) return rc;
    return run_sweeps(ctx, d_Tin, d_Tout, dt, theta, kappa, Tinf, 0, 2, 0, nullptr, nullptr, nullptr, nullptr, nullptr, st);
}

// ---- z-slab decomposition ---------------------------------------------------

int adi_cart_set_slab(adi_ctx *ctx, int rank, int nranks)
{
    int rc = check_cart(ctx, "adi_cart_set_slab");
    if (rc) return rc;
    if (nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks) {
        set_error("adi_cart_set_slab: need 0 <= rank < nranks <= 16");
        return ADI_EINVAL;
    }
    ctx->slab_rank = rank; ctx->slab_nranks = nranks;
    return ADI_OK;
}

int adi_cart_set_mask_halo(adi_ctx *ctx, const uint8_t *d_mask_lo, const uint8_t *d_mask_hi)
{
    int rc = check_cart(ctx, "adi_cart_set_mask_halo");
    if (rc) return rc;
    ctx->d_mask_lo = d_mask_lo; ctx->d_mask_hi = d_mask_hi;
    ctx->code_dirty = true;
    ctx->sparse_trust = 0;
    ctx->operand_epoch++;
    return ADI_OK;
}

int adi_cart_pack_zplanes(adi_ctx *ctx, const void *d_field, int elem_bytes, void *d_lo_out, void *d_hi_out,
                          void *stream)
{
    int rc = check_cart(ctx, "adi_cart_pack_zplanes");
    if (rc) return rc;
    if (!d_field || (elem_bytes != 1 && elem_bytes != 8)) {
        set_error("adi_cart_pack_zplanes: field is NULL or element size is not 1 / 8");
        return ADI_EINVAL;
    }
    const size_t nlines = (size_t)ctx->nx * ctx->ny;
    if (!nlines || !ctx->nz) return ADI_OK;
    const int threads = 256;
    const int blocks = (int)std::min<size_t>((nlines + threads - 1) / threads, 148 * 16);
    if (elem_bytes == 8)
        k_pack_zplanes<double><<<blocks, threads, 0, (cudaStream_t)stream>>>((const double *)d_field, (double *)d_lo_out,
                                                                         (double *)d_hi_out, nlines, ctx->nz);
    else
        k_pack_zplanes<uint8_t><<<blocks, threads, 0, (cudaStream_t)stream>>>((const uint8_t *)d_field, (uint8_t *)d_lo_out,
                                                                          (uint8_t *)d_hi_out, nlines, ctx->nz);
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    return ADI_OK;
}

int adi_cart_step_xy(adi_ctx *ctx, const double *d_Tin, double *d_Tout, const double *d_Tlo, const double *d_Thi,
                     double dt, double theta, double kappa, double Tinf, void *stream)
{
    int rc = check_cart(ctx, "adi_cart_step_xy");
    if (rc) return rc;
    if (!d_Tin || !d_Tout || d_Tin == d_Tout) {
        set_error("adi_cart_step_xy: Tin/Tout must be distinct device arrays");
        return ADI_EINVAL;
    }
    if ((ctx->d_mask_lo && !d_Tlo) || (ctx->d_mask_hi && !d_Thi)) {
        set_error("adi_cart_step_xy: a neighbouring slab is bound (mask halo) but its T plane is missing");
        return ADI_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    rc = prof_mark(ctx, 0, st);
    if (rc) return rc;
    return run_sweeps(ctx, d_Tin, d_Tout, dt, theta, kappa, Tinf, 0, 1, 0, ctx->d_mask_lo ? d_Tlo : nullptr,
                      ctx->d_mask_hi ? d_Thi : nullptr, nullptr, nullptr, nullptr, st);
}

int adi_cart_zsweep_reduce(adi_ctx *ctx, double *d_T, double *d_iface_dyn, double *d_iface_stat, double dt,
                           double theta, double kappa, double Tinf, void *stream)
{
    int rc = check_cart(ctx, "adi_cart_zsweep_reduce");
    if (rc) return rc;
    if (!d_T || !d_iface_dyn) {
        set_error("adi_cart_zsweep_reduce: NULL argument");
        return ADI_EINVAL;
    }
    return run_sweeps(ctx, d_T, d_T, dt, theta, kappa, Tinf, 2, 2, d_iface_stat ? 1 : 3, nullptr, nullptr, d_iface_dyn,
                      d_iface_stat, nullptr, (cudaStream_t)stream);
}

int adi_cart_zsweep_finish(adi_ctx *ctx, double *d_T, const double *d_dyn_all, const double *d_stat_all, double dt,
                           double theta, double kappa, double Tinf, void *stream)
{
    int rc = check_cart(ctx, "adi_cart_zsweep_finish");
    if (rc) return rc;
    if (!d_T || !d_dyn_all || !d_stat_all) {
        set_error("adi_cart_zsweep_finish: NULL argument");
        return ADI_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nlines = (size_t)ctx->nx * ctx->ny;
    if (!nlines || !ctx->nz) return ADI_OK;
    if (ctx->ghost_lines < nlines) {
        if (ctx->d_ghost) { ADI_CUDA(cudaStreamSynchronize(st)); ADI_CUDA(cudaFree(ctx->d_ghost)); }
        ctx->d_ghost = nullptr;
        ADI_CUDA(cudaMalloc(&ctx->d_ghost, 2 * nlines * sizeof(double)));
        ctx->ghost_lines = nlines;
    }
    const int threads = 128;
    const int blocks = (int)std::min<size_t>((nlines + threads - 1) / threads, 148 * 32);
    k_iface_solve<<<blocks, threads, 0, st>>>(d_dyn_all, d_stat_all, nlines, ctx->d_ghost, nlines, ctx->slab_nranks, ctx->slab_rank);
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    return run_sweeps(ctx, d_T, d_T, dt, theta, kappa, Tinf, 2, 2, 2, nullptr, nullptr, nullptr, nullptr, ctx->d_ghost, st);
}

static int ensure_ghost(adi_ctx *ctx, size_t nlines, cudaStream_t st)
{
    if (ctx->ghost_lines < nlines) {
        if (ctx->d_ghost) { ADI_CUDA(cudaStreamSynchronize(st)); ADI_CUDA(cudaFree(ctx->d_ghost)); }
        ctx->d_ghost = nullptr;
        ADI_CUDA(cudaMalloc(&ctx->d_ghost, 2 * nlines * sizeof(double)));
        ctx->ghost_lines = nlines;
    }
    return ADI_OK;
}

// the line whose unit-ghost responses stand in for every line that responds alike (k_spike_canon): mid-grid
static size_t spike_canon_line(const adi_ctx *ctx) { return (size_t)(ctx->nx / 2) * ctx->ny + (size_t)(ctx->ny / 2); }

int adi_cart_zsweep_solve0(adi_ctx *ctx, double *d_T, double *d_iface_dyn, double dt, double theta, double kappa,
                           double Tinf, void *stream)
{
    int rc = check_cart(ctx, "adi_cart_zsweep_solve0");
    if (rc) return rc;
    if (!d_T || !d_iface_dyn) {
        set_error("adi_cart_zsweep_solve0: NULL argument");
        return ADI_EINVAL;
    }
    return run_sweeps(ctx, d_T, d_T, dt, theta, kappa, Tinf, 2, 2, 4, nullptr, nullptr, d_iface_dyn, nullptr, nullptr,
                      (cudaStream_t)stream);
}

int adi_cart_zsweep_spike(adi_ctx *ctx, double *d_scratch, int end, int kmax, double threshold, double *d_compact,
                          int *d_K, int *h_maxK, double dt, double theta, double kappa, void *stream)
{
    int rc = check_cart(ctx, "adi_cart_zsweep_spike");
    if (rc) return rc;
    if (!d_scratch || !d_compact || !d_K || !h_maxK || (end != 0 && end != 1) || kmax < 1 || kmax > ctx->nz) {
        set_error("adi_cart_zsweep_spike: bad argument");
        return ADI_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nlines = (size_t)ctx->nx * ctx->ny;
    *h_maxK = 0;
    if (!nlines || !ctx->nz) return ADI_OK;
    if ((rc = ensure_ghost(ctx, nlines, st))) return rc;
    if (!ctx->d_maxk) ADI_CUDA(cudaMalloc(&ctx->d_maxk, sizeof(int)));
    ADI_CUDA(cudaMemsetAsync(d_scratch, 0, nlines * ctx->nz * sizeof(double), st));
    ADI_CUDA(cudaMemsetAsync(ctx->d_maxk, 0, sizeof(int), st));
    const int threads = 256;
    const int blocks = (int)std::min<size_t>((nlines + threads - 1) / threads, 148 * 16);
    k_fill_ghost<<<blocks, threads, 0, st>>>(ctx->d_ghost, nlines, end == 0 ? 1.0 : 0.0, end == 0 ? 0.0 : 1.0);
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    // homogeneous system: zero field, no flux, Dirichlet rows with value 0, ambient 0
    rc = run_sweeps(ctx, d_scratch, d_scratch, dt, theta, kappa, 0.0, 2, 2, 5, nullptr, nullptr, nullptr, nullptr,
                    ctx->d_ghost, st);
    if (rc) return rc;
    const int wblocks = (int)std::min<size_t>((nlines * 32 + threads - 1) / threads, 148 * 32);
    k_spike_pack<<<wblocks, threads, 0, st>>>(d_scratch, nlines, ctx->nz, end, kmax, threshold, d_compact, d_K,
                                              ctx->d_maxk);
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    k_spike_canon<<<blocks, threads, 0, st>>>(d_compact, d_K, nlines, kmax, spike_canon_line(ctx));
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    ADI_CUDA(cudaMemcpyAsync(h_maxK, ctx->d_maxk, sizeof(int), cudaMemcpyDeviceToHost, st));
    ADI_CUDA(cudaStreamSynchronize(st));
    return ADI_OK;
}

int adi_cart_zsweep_apply(adi_ctx *ctx, double *d_T, const double *d_dyn_all, const double *d_stat_all,
                          const double *d_vC, const double *d_wC, const int *d_Kv, const int *d_Kw, int kmax,
                          void *stream)
{
    int rc = check_cart(ctx, "adi_cart_zsweep_apply");
    if (rc) return rc;
    if (!d_T || !d_dyn_all || !d_stat_all || !d_vC || !d_wC || !d_Kv || !d_Kw || kmax < 1 || kmax > ctx->nz) {
        set_error("adi_cart_zsweep_apply: bad argument");
        return ADI_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nlines = (size_t)ctx->nx * ctx->ny;
    if (!nlines || !ctx->nz) return ADI_OK;
    if ((rc = ensure_ghost(ctx, nlines, st))) return rc;
    const int threads = 128;
    const int blocks = (int)std::min<size_t>((nlines + threads - 1) / threads, 148 * 32);
    k_iface_solve<<<blocks, threads, 0, st>>>(d_dyn_all, d_stat_all, nlines, ctx->d_ghost, nlines, ctx->slab_nranks, ctx->slab_rank);
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    const int ablocks = (int)std::min<size_t>((nlines * 8 + 255) / 256, 148 * 32);
    const size_t cl = spike_canon_line(ctx) * (size_t)kmax;
    k_spike_apply<false><<<ablocks, 256, 0, st>>>(d_T, ctx->d_ghost, d_vC, d_wC, d_Kv, d_Kw, nlines, ctx->nz, kmax, d_vC + cl,
                                                  d_wC + cl);
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    return prof_mark(ctx, 4, st);
}

int adi_cart_step_host(adi_ctx *ctx, const double *h_Tin, double *h_Tout, int nsteps, double dt,
                       double theta, double kappa, double Tinf, void *stream)
{
    int rc = check_cart(ctx, "adi_cart_step_host");
    if (rc) return rc;
    if (!h_Tin || !h_Tout || nsteps < 1) {
        set_error("adi_cart_step_host: bad arguments");
        return ADI_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ncell = (size_t)ctx->nx * ctx->ny * ctx->nz;
    rc = ensure_stage(ctx, ncell);
    if (rc) return rc;
    if ((rc = stage_h2d(ctx, ctx->stage[0], h_Tin, ncell * sizeof(double), st))) return rc;
    int cur = 0;
    for (int s = 0; s < nsteps; ++s) {
        rc = adi_cart_step(ctx, ctx->stage[cur], ctx->stage[cur ^ 1], dt, theta, kappa, Tinf, stream);
        if (rc) return rc;
        cur ^= 1;
    }
    return stage_d2h(ctx, h_Tout, ctx->stage[cur], ncell * sizeof(double), st);
}

int adi_cart_step_host_async(adi_ctx *ctx, int slot, const double *h_Tin, double *h_Tout, double dt,
                             double theta, double kappa, double Tinf, void *stream)
{
    int rc = check_cart(ctx, "adi_cart_step_host_async");
    if (rc) return rc;
    if (!h_Tin || !h_Tout || slot < 0 || slot > 1) {
        set_error("adi_cart_step_host_async: bad arguments (slot must be 0 or 1)");
        return ADI_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t ncell = (size_t)ctx->nx * ctx->ny * ctx->nz;
    if (ctx->pipe_cells[slot] < ncell || !ctx->pipe[slot][0]) {
        for (int j = 0; j < 2; ++j) {
            if (ctx->pipe[slot][j]) { ADI_CUDA(cudaDeviceSynchronize()); ADI_CUDA(cudaFree(ctx->pipe[slot][j])); }
            ctx->pipe[slot][j] = nullptr;
            ADI_CUDA(cudaMalloc(&ctx->pipe[slot][j], std::max<size_t>(ncell, 1) * sizeof(double)));
        }
        ctx->pipe_cells[slot] = ncell;
    }
    ADI_CUDA(cudaMemcpyAsync(ctx->pipe[slot][0], h_Tin, ncell * sizeof(double), cudaMemcpyHostToDevice, st));
    rc = adi_cart_step(ctx, ctx->pipe[slot][0], ctx->pipe[slot][1], dt, theta, kappa, Tinf, stream);
    if (rc) return rc;
    ADI_CUDA(cudaMemcpyAsync(h_Tout, ctx->pipe[slot][1], ncell * sizeof(double), cudaMemcpyDeviceToHost, st));
    return ADI_OK;
}

int adi_cart_build_packs(adi_ctx *ctx, double rho, double cp, const int h_kind[6],
                         const double h_scalar[6], const double *const d_h_field[6],
                         const int q_kind[6], const double q_scalar[6],
                         const double *const d_q_field[6], double *d_coeff_x, double *d_coeff_y,
                         double *d_coeff_z, double *d_q_x, double *d_q_y, double *d_q_z,
                         void *stream)
{
    int rc = check_cart(ctx, "adi_cart_build_packs");
    if (rc) return rc;
    if (!ctx->d_mask) {
        set_error("adi_cart_build_packs: no mask bound");
        return ADI_ESTATE;
    }
    PackArgs a;
    a.mask = ctx->d_mask;
    a.mlo = ctx->d_mask_lo; a.mhi = ctx->d_mask_hi;
    a.nx = ctx->nx; a.ny = ctx->ny; a.nz = ctx->nz;
    const double dx = ctx->dx;
    a.A = dx * dx;
    const double V = std::pow(dx, 3.0);  // dx**3 (adi3d_numba_coeff.py:70): Python float power = C pow()
    a.Ccell = rho * cp * V;
    for (int f = 0; f < 6; ++f) {
        a.h_kind[f] = h_kind ? h_kind[f] : 0;
        a.h_scalar[f] = h_scalar ? h_scalar[f] : 0.0;
        a.h_field[f] = d_h_field ? d_h_field[f] : nullptr;
        a.q_kind[f] = q_kind ? q_kind[f] : 0;
        a.q_scalar[f] = q_scalar ? q_scalar[f] : 0.0;
        a.q_field[f] = d_q_field ? d_q_field[f] : nullptr;
        if ((a.h_kind[f] == 2 && !a.h_field[f]) || (a.q_kind[f] == 2 && !a.q_field[f]) ||
            a.h_kind[f] < 0 || a.h_kind[f] > 2 || a.q_kind[f] < 0 || a.q_kind[f] > 2) {
            set_error("adi_cart_build_packs: bad kind / missing field");
            return ADI_EINVAL;
        }
    }
    a.coeff[0] = d_coeff_x; a.coeff[1] = d_coeff_y; a.coeff[2] = d_coeff_z;
    a.qout[0] = d_q_x; a.qout[1] = d_q_y; a.qout[2] = d_q_z;
    const size_t n = (size_t)a.nx * a.ny * a.nz;
    if (!n) return ADI_OK;
    const int threads = 256;
    const int blocks = (int)std::min<size_t>((n + threads - 1) / threads, 148 * 64);
    // word form (adi_mask_core.h): 2 cells per thread when z lines are 2-byte aligned and the outputs take 16-byte stores.
    // Measured at 1024^3, half-built part (profiles/r04f_packs_probe.txt): 4.6 ms = 5.6 TB/s of stores (memset of the
    // three fields: 3.4 ms) against 7.0 ms for the cell form; 40 registers (6 blocks per SM), streaming stores, 512
    // blocks per SM.  Option pkm: 1 = plain stores, 4 = 4 cells per thread (half-sector stores, 7.9 ms); pkb: blocks per SM.
    uintptr_t al = 0;
    for (int ax = 0; ax < 3; ++ax) al |= (uintptr_t)a.coeff[ax] | (uintptr_t)a.qout[ax];
    ctx->maskv_used &= ~4;
    const int nc = (ctx->opt_pkm & 4) ? 4 : 2;
    if (ctx->opt_maskv && a.nz % nc == 0 && ((uintptr_t)a.mask & (nc - 1)) == 0 && (al & 15) == 0) {
        const int per_sm = ctx->opt_pkb > 0 ? ctx->opt_pkb : 512;
        const int vblocks = (int)std::min<size_t>((n / nc + threads - 1) / threads, (size_t)148 * per_sm);
        cudaStream_t st = (cudaStream_t)stream;
        if (nc == 4) k_build_packs_v<4, true, 1><<<vblocks, threads, 0, st>>>(a);
        else if (ctx->opt_pkm & 1) k_build_packs_v<2, false, 6><<<vblocks, threads, 0, st>>>(a);
        else k_build_packs_v<2, true, 6><<<vblocks, threads, 0, st>>>(a);
        ctx->maskv_used |= 4;
    } else {
        k_build_packs<<<blocks, threads, 0, (cudaStream_t)stream>>>(a);
    }
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    return ADI_OK;
}

int adi_cart_exposed_mask(adi_ctx *ctx, int face, uint8_t *d_out, void *stream)
{
    int rc = check_cart(ctx, "adi_cart_exposed_mask");
    if (rc) return rc;
    if (face < 0 || face > 5) {
        set_error("bad face");  // ValueError("bad face") adi3d_gpu_coeff.py:47
        return ADI_EINVAL;
    }
    if (!ctx->d_mask || !d_out) {
        set_error("adi_cart_exposed_mask: no mask bound / no output");
        return ADI_ESTATE;
    }
    const size_t n = (size_t)ctx->nx * ctx->ny * ctx->nz;
    if (!n) return ADI_OK;
    const int threads = 256;
    const int blocks = (int)std::min<size_t>((n + threads - 1) / threads, 148 * 64);
    k_exposed_mask<<<blocks, threads, 0, (cudaStream_t)stream>>>(ctx->d_mask, d_out, face, ctx->nx,
                                                                ctx->ny, ctx->nz, ctx->d_mask_lo, ctx->d_mask_hi);
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    return ADI_OK;
}

}  // extern "C"

// ---- internal entry points of the in-library multi-GPU sequencing (adi_dist.cu) ---------------------------
namespace adi {

int cart_ensure_ghost(adi_ctx *ctx, size_t nlines, cudaStream_t st) { return ensure_ghost(ctx, nlines, st); }

// "solve first" z pass of the lines [line0, line0 + nlb): in place with zero ghosts, (y_first, y_last) -> d_dyn[2][nlb]
int cart_zsolve0_range(adi_ctx *ctx, double *d_T, double *d_dyn, double dt, double theta, double kappa, double Tinf,
                       size_t line0, size_t nlb, cudaStream_t st)
{
    return run_sweeps(ctx, d_T, d_T, dt, theta, kappa, Tinf, 2, 2, 4, nullptr, nullptr, d_dyn, nullptr, nullptr, st, line0, nlb);
}

// inter-rank solve + ghost corrections of the lines [line0, line0 + nlb): d_dyn_all[R][2][nlb] is the batch's gathered
// right-hand-side part, d_stat_all[R][4][nl] the matrix part of all lines
int cart_zapply_range(adi_ctx *ctx, double *d_T, const double *d_dyn_all, const double *d_stat_all, const double *d_vC,
                      const double *d_wC, const int *d_Kv, const int *d_Kw, int kmax, size_t line0, size_t nlb,
                      cudaStream_t st)
{
    const size_t nl = (size_t)ctx->nx * ctx->ny;
    const int ablocks = (int)std::min<size_t>((nlb * 8 + 255) / 256, 148 * 32);
    const size_t cl = spike_canon_line(ctx) * (size_t)kmax;
    k_spike_apply<true><<<ablocks, 256, 0, st>>>(d_T + line0 * (size_t)ctx->nz, nullptr, d_vC + line0 * (size_t)kmax,
                                                 d_wC + line0 * (size_t)kmax, d_Kv + line0, d_Kw + line0, nlb, ctx->nz, kmax,
                                                 d_vC + cl, d_wC + cl, d_dyn_all, d_stat_all + line0, nl, ctx->slab_nranks,
                                                 ctx->slab_rank);
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    return ADI_OK;
}

int cart_prof_mark(adi_ctx *ctx, int slot, cudaStream_t st) { return prof_mark(ctx, slot, st); }

// Neighbour code and operand checks up to date (on `st`): what the deferred explicit stage needs before another
// stream may run k_explicit_faces.
int cart_prepare(adi_ctx *ctx, cudaStream_t st)
{
    int rc = ensure_code(ctx, st);
    if (rc) return rc;
    return ensure_sparse(ctx, st);
}

// Explicit stage of the slab-face cells whose stencil reaches into the adjacent slabs (k_explicit_faces), queued on
// `st` (the communication stream, behind the arrival of the planes).  No-op without an explicit stage.
int cart_explicit_faces(adi_ctx *ctx, const double *d_Tin, double *d_Tout, const double *d_Tlo, const double *d_Thi,
                        double dt, double theta, double kappa, cudaStream_t st)
{
    const size_t nlines = (size_t)ctx->nx * ctx->ny;
    if (!nlines || !ctx->nz || theta == 1.0) return ADI_OK;
    const double dx = ctx->dx;
    SweepArgs a = {};
    a.nx = ctx->nx; a.ny = ctx->ny; a.nz = ctx->nz;
    a.k.g = theta * (kappa * dt / (dx * dx));
    a.k.dt = dt;
    a.k.beta = dt * kappa * (1.0 - theta);
    a.k.invdx2 = 1.0 / (dx * dx);
    a.in = d_Tin; a.out = d_Tout; a.code = ctx->code[0];
    a.zlo = ctx->d_mask_lo ? d_Tlo : nullptr; a.zhi = ctx->d_mask_hi ? d_Thi : nullptr;
    if (!a.zlo && !a.zhi) return ADI_OK;
    k_explicit_faces<<<(unsigned)std::min<size_t>((nlines + 127) / 128, 148 * 32), 128, 0, st>>>(a);
    ctx->launches++;
    ADI_CUDA(cudaGetLastError());
    return ADI_OK;
}

// adi_cart_step_xy with the T planes of the adjacent slabs still on their way: `halo_ready` is recorded (on another
// stream) behind cart_explicit_faces
int cart_step_xy_deferred(adi_ctx *ctx, const double *d_Tin, double *d_Tout, const double *d_Tlo, const double *d_Thi,
                          double dt, double theta, double kappa, double Tinf, cudaStream_t st, cudaEvent_t halo_ready)
{
    int rc = prof_mark(ctx, 0, st);
    if (rc) return rc;
    return run_sweeps(ctx, d_Tin, d_Tout, dt, theta, kappa, Tinf, 0, 1, 0, ctx->d_mask_lo ? d_Tlo : nullptr,
                      ctx->d_mask_hi ? d_Thi : nullptr, nullptr, nullptr, nullptr, st, 0, 0, halo_ready);
}

}  // namespace adi
