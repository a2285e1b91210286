// adi_api.cu -- context, memory and error plumbing of the C ABI (include/adi_b200.h).
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <new>
#include <thread>
#include <vector>

#include "adi_ctx.h"

namespace adi {

static thread_local std::string g_err;

void set_error(const std::string &msg) { g_err = msg; }

int cuda_fail(cudaError_t e, const char *what)
{
    g_err = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return ADI_ECUDA;
}

void cyl_release(adi_ctx *ctx);  // adi_cyl.cu

int prof_mark(adi_ctx *ctx, int slot, cudaStream_t st)
{
    if (!ctx->opt_profile) return ADI_OK;
    const size_t need = (size_t)(ctx->prof_steps + 1) * 5;
    while (ctx->prof_ev.size() < need) {
        cudaEvent_t e;
        ADI_CUDA(cudaEventCreate(&e));
        ctx->prof_ev.push_back(e);
    }
    ADI_CUDA(cudaEventRecord(ctx->prof_ev[(size_t)ctx->prof_steps * 5 + slot], st));
    if (slot == 4) ctx->prof_steps++;
    return ADI_OK;
}

// ---- staged copies of pageable host arrays ------------------------------------------------------------
static const size_t PIN_CHUNK = (size_t)32 << 20;

static bool host_is_pinned(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

static int ensure_pin(adi_ctx *ctx)
{
    for (int i = 0; i < 2; ++i) {
        if (!ctx->pin[i]) ADI_CUDA(cudaHostAlloc(&ctx->pin[i], PIN_CHUNK, cudaHostAllocDefault));
        if (!ctx->pin_ev[i]) ADI_CUDA(cudaEventCreateWithFlags(&ctx->pin_ev[i], cudaEventDisableTiming));
    }
    return ADI_OK;
}

static void par_memcpy(void *dst, const void *src, size_t n)
{
    unsigned hw = std::thread::hardware_concurrency();
    const size_t nt = std::min<size_t>(std::max(1u, std::min(hw, 8u)), (n + ((size_t)4 << 20) - 1) / ((size_t)4 << 20));
    if (nt <= 1) {
        memcpy(dst, src, n);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((n + nt - 1) / nt + 63) & ~(size_t)63;
    for (size_t t = 0; t < nt; ++t) {
        const size_t o = t * per;
        if (o >= n) break;
        const size_t m = std::min(per, n - o);
        th.emplace_back([=] { memcpy((char *)dst + o, (const char *)src + o, m); });
    }
    for (auto &t : th) t.join();
}

int stage_h2d(adi_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, cudaStream_t st)
{
    if (!bytes) return ADI_OK;
    if (bytes < ((size_t)1 << 20) || host_is_pinned(h_src)) {
        ADI_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
        return ADI_OK;
    }
    int rc = ensure_pin(ctx);
    if (rc) return rc;
    size_t off = 0;
    for (int i = 0; off < bytes; ++i, off += PIN_CHUNK) {
        const int b = i & 1;
        const size_t n = std::min(PIN_CHUNK, bytes - off);
        if (i >= 2) ADI_CUDA(cudaEventSynchronize(ctx->pin_ev[b]));   // the transfer out of this buffer is done
        par_memcpy(ctx->pin[b], (const char *)h_src + off, n);
        ADI_CUDA(cudaMemcpyAsync((char *)d_dst + off, ctx->pin[b], n, cudaMemcpyHostToDevice, st));
        ADI_CUDA(cudaEventRecord(ctx->pin_ev[b], st));
    }
    // the staging buffers may be reused by the next call straight away: wait for the last two transfers
    ADI_CUDA(cudaEventSynchronize(ctx->pin_ev[0]));
    ADI_CUDA(cudaEventSynchronize(ctx->pin_ev[1]));
    return ADI_OK;
}

int stage_d2h(adi_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, cudaStream_t st)
{
    if (!bytes) return ADI_OK;
    if (bytes < ((size_t)1 << 20) || host_is_pinned(h_dst)) {
        ADI_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
        ADI_CUDA(cudaStreamSynchronize(st));
        return ADI_OK;
    }
    int rc = ensure_pin(ctx);
    if (rc) return rc;
    const size_t nchunk = (bytes + PIN_CHUNK - 1) / PIN_CHUNK;
    for (size_t i = 0; i <= nchunk; ++i) {
        if (i < nchunk) {
            const int b = (int)(i & 1);
            const size_t off = i * PIN_CHUNK, n = std::min(PIN_CHUNK, bytes - off);
            ADI_CUDA(cudaMemcpyAsync(ctx->pin[b], (const char *)d_src + off, n, cudaMemcpyDeviceToHost, st));
            ADI_CUDA(cudaEventRecord(ctx->pin_ev[b], st));
        }
        if (i >= 1) {   // piece i-1 has had the time of one transfer to arrive
            const int b = (int)((i - 1) & 1);
            const size_t off = (i - 1) * PIN_CHUNK, n = std::min(PIN_CHUNK, bytes - off);
            ADI_CUDA(cudaEventSynchronize(ctx->pin_ev[b]));
            par_memcpy((char *)h_dst + off, ctx->pin[b], n);
        }
    }
    return ADI_OK;
}

void stage_release(adi_ctx *ctx)
{
    for (int i = 0; i < 2; ++i) {
        if (ctx->pin[i]) cudaFreeHost(ctx->pin[i]);
        if (ctx->pin_ev[i]) cudaEventDestroy(ctx->pin_ev[i]);
        ctx->pin[i] = nullptr; ctx->pin_ev[i] = nullptr;
    }
}

}  // namespace adi

extern "C" {

const char *adi_last_error(void) { return adi::g_err.c_str(); }

const char *adi_version(void) { return "adi_b200 0.1 (sm_100a)"; }

int adi_ctx_create(int device, adi_ctx **out)
{
    if (!out) {
        adi::set_error("adi_ctx_create: out is NULL");
        return ADI_EINVAL;
    }
    int ndev = 0;
    ADI_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) {
        adi::set_error("adi_ctx_create: no such CUDA device");
        return ADI_EINVAL;
    }
    ADI_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ADI_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        adi::set_error("adi_ctx_create: this library is built for sm_100a (B200) only");
        return ADI_EINVAL;
    }
    adi_ctx *c = new (std::nothrow) adi_ctx();
    if (!c) return ADI_ENOMEM;
    c->device = device;
    *out = c;
    return ADI_OK;
}

int adi_ctx_destroy(adi_ctx *ctx)
{
    if (!ctx) return ADI_OK;
    cudaSetDevice(ctx->device);
    for (int a = 0; a < 3; ++a)
        if (ctx->code_buf[a]) cudaFree(ctx->code_buf[a]);
    for (int a = 0; a < 2; ++a)
        if (ctx->codeT[a]) cudaFree(ctx->codeT[a]);
    for (int a = 0; a < 3; ++a)
        if (ctx->tiles[a].d) { cudaFree(ctx->tiles[a].d); cudaFree(ctx->tiles[a].d_uni); cudaFree(ctx->tiles[a].d_gen); }
    if (ctx->d_tflags) cudaFree(ctx->d_tflags);
    if (ctx->h_tflags) cudaFreeHost(ctx->h_tflags);
    for (int a = 0; a < 2; ++a)
        if (ctx->stage[a]) cudaFree(ctx->stage[a]);
    if (ctx->d_ghost) cudaFree(ctx->d_ghost);
    if (ctx->d_maxk) cudaFree(ctx->d_maxk);
    if (ctx->d_viol) cudaFree(ctx->d_viol);
    if (ctx->d_ztop) cudaFree(ctx->d_ztop);
    if (ctx->h_ztop) cudaFreeHost(ctx->h_ztop);
    if (ctx->h_viol) cudaFreeHost(ctx->h_viol);
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            if (ctx->pipe[i][j]) cudaFree(ctx->pipe[i][j]);
    if (ctx->stage_mask) cudaFree(ctx->stage_mask);
    if (ctx->stage_src) cudaFree(ctx->stage_src);
    adi::cyl_release(ctx);
    adi::text_release(ctx);
    adi::stage_release(ctx);
    adi::dist_release(ctx);
    for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
    delete ctx;
    return ADI_OK;
}

int adi_sync(adi_ctx *ctx, void *stream)
{
    if (!ctx) return ADI_EINVAL;
    ADI_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return ADI_OK;
}

int adi_malloc(adi_ctx *ctx, size_t bytes, void **d_ptr)
{
    if (!ctx || !d_ptr) return ADI_EINVAL;
    ADI_CUDA(cudaSetDevice(ctx->device));
    ADI_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 1));
    return ADI_OK;
}

int adi_free(adi_ctx *ctx, void *d_ptr)
{
    if (!ctx) return ADI_EINVAL;
    ADI_CUDA(cudaFree(d_ptr));
    return ADI_OK;
}

int adi_h2d(adi_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, void *stream)
{
    if (!ctx) return ADI_EINVAL;
    ADI_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return ADI_OK;
}

int adi_d2h(adi_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, void *stream)
{
    if (!ctx) return ADI_EINVAL;
    ADI_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    ADI_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return ADI_OK;
}

int adi_set_option(adi_ctx *ctx, const char *name, long value)
{
    if (!ctx || !name) return ADI_EINVAL;
    if (!strcmp(name, "kt")) ctx->opt_kt = value;
    else if (!strcmp(name, "lt")) ctx->opt_lt = value;
    else if (!strcmp(name, "m")) ctx->opt_m = value;
    else if (!strcmp(name, "sync_check")) ctx->opt_sync_check = value;
    else if (!strcmp(name, "profile")) ctx->opt_profile = value;
    else if (!strcmp(name, "wide")) ctx->opt_wide = value;  // 1: 512-thread blocks for lines <= 512 cells too
    else if (!strcmp(name, "sparse_coeff")) {  // 1 (default): read verified surface-only coefficient fields at exposed cells only
        ctx->opt_sparse = value;
        ctx->sparse_dirty = true;
    }
    else if (!strcmp(name, "sparse_trust")) {
        ctx->sparse_trust = value;
        ctx->sparse_dirty = true;
    }
    else if (!strcmp(name, "xy2")) {   // 1 (default): second-generation x / y sweeps (adi_sweep_xy.cuh)
        ctx->opt_xy2 = value;
        ctx->code_dirty = true;
    }
    else if (!strcmp(name, "uni")) ctx->opt_uni = value;    // 1 (default): tabulated factors for uniform chunks
    else if (!strcmp(name, "tw")) ctx->opt_tw = value;      // 1: reduced system solved by warps (default 0: shared-memory PCR, measured faster)
    else if (!strcmp(name, "zt")) {     // 1 (default): second-generation z sweep (adi_sweep_zt.cuh)
        ctx->opt_zt = value;
        ctx->sparse_dirty = true;
    }
    else if (!strcmp(name, "tiles")) {  // 1 (default): sweeps launch only the tiles that hold an active cell
        ctx->opt_tiles = value;
        for (int a = 0; a < 3; ++a) ctx->tiles[a].valid = false;
    }
    else if (!strcmp(name, "xyu")) ctx->opt_xyu = value;      // 1 (default): all-uniform tiles of 1025..2048-cell x / y lines on k_sweep_xyu (two blocks per SM)
    else if (!strcmp(name, "cylsm")) ctx->opt_cylsm = value;  // cylindrical strided sweeps: 1 (default) tables staged in shared memory, blocks walk over 4 z tiles (n > 1: n tiles); 0 tables through L1
    else if (!strcmp(name, "cylzt")) ctx->opt_cylzt = value;  // 1 (default): unmasked cylindrical z sweep on the Cartesian z kernel (k_sweep_zt)
    else if (!strcmp(name, "maskv")) {  // 1 (default): word forms of the per-mask-change kernels (adi_mask_core.h) where alignment allows
        ctx->opt_maskv = value;
        ctx->code_dirty = true;
    }
    else if (!strcmp(name, "ztrim")) ctx->opt_ztrim = value;  // 1 (default): the z sweep stops at the top of a part under construction
    else if (!strcmp(name, "pkb")) ctx->opt_pkb = value;      // word-form pack builder: blocks per SM (0: default 512)
    else if (!strcmp(name, "pkm")) ctx->opt_pkm = value;      // word-form pack builder tuning aids: 1 plain instead of streaming stores, 4 four cells per thread
    else if (!strcmp(name, "ukt")) ctx->opt_ukt = value;
    else if (!strcmp(name, "zm")) ctx->opt_zm = value;        // k_sweep_zt: chunk length (16 / 32) whatever the line length
    else if (!strcmp(name, "ejt")) ctx->opt_ejt = value;      // explicit stage: y rows per block (default 16)
    else if (!strcmp(name, "eth")) ctx->opt_eth = value;      // explicit stage: threads per block (default 128)
    else if (!strcmp(name, "eorder")) ctx->opt_eorder = value;  // explicit stage: 1 = blocks of neighbouring x planes run together
    else if (!strcmp(name, "hyb")) ctx->opt_hyb = value;    // 1 (default): z sweep, uniform lead + general tail in one chunk (the chunk under a surface)
    else if (!strcmp(name, "xyp")) ctx->opt_xyp = value;    // 1: x / y lines of 1025..2048 cells on persistent blocks, tiles prefetched by the TMA engine (default 0: no faster than k_sweep_xy, r02o-r02q)
    else if (!strcmp(name, "seq")) ctx->opt_seq = value;    // k_sweep_xyp: 1 contiguous tile range per block, 0 (default) round-robin
    else if (!strcmp(name, "promo")) ctx->opt_promo = value;  // k_sweep_xyp: L2 promotion of the tensor map (0 / 64 / 128 / 256 bytes)
    else if (!strcmp(name, "lb")) ctx->opt_lb = value;      // 256: x / y lines of 1025..2048 cells in 256-thread blocks
    else if (!strcmp(name, "bulk")) ctx->opt_bulk = value;  // 1 (default): z sweep tiles as bulk asynchronous copies
    else if (!strcmp(name, "occ")) ctx->opt_occ = value;    // x / y sweeps, 16-cell chunks: resident blocks per SM (2, 3, 4)
    else if (!strcmp(name, "remap")) ctx->opt_remap = value;  // 1: both ends of a line in one warp (measured slower)
    else if (!strcmp(name, "dbg")) ctx->opt_dbg = value;    // tuning aid (1: x / y sweeps move data only -- wrong results)
    else if (!strcmp(name, "fuse")) ctx->opt_fuse = value;  // 1: explicit stage fused into the x sweep
    else {
        adi::set_error(std::string("adi_set_option: unknown option ") + name);
        return ADI_EINVAL;
    }
    return ADI_OK;
}

long adi_get_option(adi_ctx *ctx, const char *name)
{
    if (!ctx || !name) return -1;
    if (!strcmp(name, "kt")) return ctx->opt_kt;
    if (!strcmp(name, "lt")) return ctx->opt_lt;
    if (!strcmp(name, "m")) return ctx->opt_m;
    if (!strcmp(name, "sync_check")) return ctx->opt_sync_check;
    if (!strcmp(name, "profile")) return ctx->opt_profile;
    if (!strcmp(name, "wide")) return ctx->opt_wide;
    if (!strcmp(name, "fuse")) return ctx->opt_fuse;
    if (!strcmp(name, "xy2")) return ctx->opt_xy2;
    if (!strcmp(name, "uni")) return ctx->opt_uni;
    if (!strcmp(name, "tw")) return ctx->opt_tw;
    if (!strcmp(name, "remap")) return ctx->opt_remap;
    if (!strcmp(name, "occ")) return ctx->opt_occ;
    if (!strcmp(name, "zt")) return ctx->opt_zt;
    if (!strcmp(name, "hyb")) return ctx->opt_hyb;
    if (!strcmp(name, "xyp")) return ctx->opt_xyp;
    if (!strcmp(name, "xyp_used")) return ctx->xyp_used;     // launches of k_sweep_xyp so far
    if (!strcmp(name, "xyp_layout")) return ctx->xyp_state;  // tensor-map layout in use (0 / 1), -1: refused by the driver
    if (!strcmp(name, "seq")) return ctx->opt_seq;
    if (!strcmp(name, "promo")) return ctx->opt_promo;
    if (!strcmp(name, "lb")) return ctx->opt_lb;
    if (!strcmp(name, "bulk")) return ctx->opt_bulk;
    if (!strcmp(name, "xyu")) return ctx->opt_xyu;
    if (!strcmp(name, "xyu_used")) return ctx->xyu_used;     // launches of k_sweep_xyu so far
    if (!strcmp(name, "cylsm")) return ctx->opt_cylsm;
    if (!strcmp(name, "cylzt")) return ctx->opt_cylzt;
    if (!strcmp(name, "ukt")) return ctx->opt_ukt;
    if (!strcmp(name, "maskv")) return ctx->opt_maskv;
    if (!strcmp(name, "ztrim")) return ctx->opt_ztrim;
    if (!strcmp(name, "ztrim_used")) return ctx->ztrim_used;   // z sweeps launched on trimmed lines so far
    if (!strcmp(name, "xtrim_used")) return ctx->xtrim_used;   // x sweeps launched on the x extent of the part only
    if (!strcmp(name, "xlo")) return ctx->xlo;
    if (!strcmp(name, "xhi")) return ctx->xhi;
    if (!strcmp(name, "ztop")) return ctx->ztop;               // z + 1 of the highest active cell (-1: unknown / not read yet)
    if (!strcmp(name, "pkb")) return ctx->opt_pkb;
    if (!strcmp(name, "pkm")) return ctx->opt_pkm;
    if (!strcmp(name, "maskv_used")) return ctx->maskv_used;   // bit 0 / 1 / 2: the last code build / transposes / pack build took the word form
    if (!strcmp(name, "zm")) return ctx->opt_zm;
    if (!strcmp(name, "ejt")) return ctx->opt_ejt;
    if (!strcmp(name, "eth")) return ctx->opt_eth;
    if (!strcmp(name, "eorder")) return ctx->opt_eorder;
    if (!strcmp(name, "tiles")) return ctx->opt_tiles;
    if (!strcmp(name, "tiles_active")) return ctx->tiles[0].n + ctx->tiles[1].n + ctx->tiles[2].n;
    if (!strcmp(name, "tiles_total")) return ctx->tiles[0].total + ctx->tiles[1].total + ctx->tiles[2].total;
    if (!strcmp(name, "dbg")) return ctx->opt_dbg;
    if (!strcmp(name, "sparse_coeff")) return ctx->opt_sparse;
    if (!strcmp(name, "sparse_active"))  // bit a: the sweep along axis a currently skips interior coefficient reads
        return ctx->sparse_dirty ? 0 : ((ctx->sparse[0] ? 1 : 0) | (ctx->sparse[1] ? 2 : 0) | (ctx->sparse[2] ? 4 : 0));
    adi::set_error(std::string("adi_get_option: unknown option ") + name);
    return -1;
}

long adi_launch_count(adi_ctx *ctx) { return ctx ? ctx->launches + adi::text_launches(ctx) : -1; }

int adi_profile_reset(adi_ctx *ctx)
{
    if (!ctx) return ADI_EINVAL;
    ctx->prof_steps = 0;
    return ADI_OK;
}

int adi_profile_read(adi_ctx *ctx, double ms[4], long *nsteps)
{
    if (!ctx || !ms) return ADI_EINVAL;
    ms[0] = ms[1] = ms[2] = ms[3] = 0.0;
    for (long s = 0; s < ctx->prof_steps; ++s) {
        cudaEvent_t *e = &ctx->prof_ev[(size_t)s * 5];
        ADI_CUDA(cudaEventSynchronize(e[4]));
        for (int k = 0; k < 4; ++k) {
            float t = 0.f;
            ADI_CUDA(cudaEventElapsedTime(&t, e[k], e[k + 1]));
            ms[k] += (double)t;
        }
    }
    if (nsteps) *nsteps = ctx->prof_steps;
    return ADI_OK;
}

}  // extern "C"
