// adi_ctx.h -- the opaque context behind include/adi_b200.h (host side).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/adi_b200.h"

namespace adi {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);

#define ADI_CUDA(call)                                            \
    do {                                                          \
        cudaError_t _e = (call);                                  \
        if (_e != cudaSuccess) return adi::cuda_fail(_e, #call);  \
    } while (0)

struct Pack {
    const double *coeff = nullptr;
    const uint8_t *dirm = nullptr;
    const double *dirv = nullptr;
    const double *q = nullptr;
};

// Active-tile list of one sweep axis (adi_cart.cu ensure_tiles): built lazily for the tile height the launcher
// uses, dropped whenever the neighbour code is rebuilt.
struct TileList {
    int *d = nullptr;        // device: ids of the tiles with an active cell, ascending
    size_t cap = 0;
    int n = 0, total = 0, kt = 0;
    bool valid = false;
    // x / y sweeps: the active tiles split into ALL-UNIFORM ones (k_tile_flags bit 1: k_sweep_xyu) and the rest
    int *d_uni = nullptr, *d_gen = nullptr;
    int n_uni = 0, n_gen = 0;
};
// -> *list = device list (NULL when every tile is active or the option is off), *nactive = tiles to launch
int ensure_tiles(adi_ctx *ctx, int axis, int KT, cudaStream_t st, const int **list, int *nactive, int *tiles_nx);
// resolves ctx->ztop / xlo / xhi (the extent of the part, reduced by the code build) after a mask change
int part_extent(adi_ctx *ctx, cudaStream_t st);
// x / y axes: the same tiles as two lists, all-uniform tiles and the other active ones (either may be empty)
int ensure_tiles_split(adi_ctx *ctx, int axis, int KT, cudaStream_t st, const int **uni, int *nuni, const int **gen, int *ngen,
                       int *tiles_nx);

struct CylTables;  // adi_cyl.cu
struct DistState;  // adi_dist.cu: NCCL communicator, exchange buffers and caches of the in-library z-slab step
void dist_release(adi_ctx *ctx);
struct TextState;  // adi_text.cu: scratch of the ASCII output path and the probe slots
void text_release(adi_ctx *ctx);
long text_launches(adi_ctx *ctx);

// profile helpers (adi_api.cu): record event #slot (0..4) of the current step:
// 0 start, 1 after the explicit stage, 2 after the x|r sweep, 3 after y|phi, 4 after z
int prof_mark(adi_ctx *ctx, int slot, cudaStream_t st);

// Host <-> device copies of the host-array entry points (adi_api.cu).  Page-locked host memory goes straight to
// cudaMemcpyAsync; pageable memory is moved through two page-locked staging buffers in 32 MiB pieces, the host-side
// memcpy of piece i+1 (several threads) overlapping the PCIe transfer of piece i -- the driver's own pageable path
// runs at ~6.5 GB/s, this one at the speed of the host memcpy.  d2h returns with the data in h_dst.
int stage_h2d(adi_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, cudaStream_t st);
int stage_d2h(adi_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, cudaStream_t st);
void stage_release(adi_ctx *ctx);

}  // namespace adi

struct adi_ctx {
    int device = 0;
    long launches = 0;
    // options (adi_set_option)
    long opt_kt = 0, opt_lt = 0, opt_m = 0, opt_sync_check = 0, opt_profile = 0, opt_fuse = 0, opt_wide = 0, opt_sparse = 1,
         opt_xy2 = 1, opt_uni = 1, opt_tw = 0, opt_remap = 0, opt_dbg = 0, opt_occ = 0, opt_zt = 1, opt_bulk = 1, opt_tiles = 1, opt_eorder = 0, opt_lb = 0, opt_hyb = 1, opt_xyp = 0, opt_seq = 0, opt_promo = 0, opt_ejt = 0, opt_eth = 0, opt_zm = 0, opt_xyu = 1, opt_ukt = 0, opt_cylsm = 1, opt_cylzt = 1, opt_maskv = 1, opt_pkb = 0, opt_pkm = 0, opt_ztrim = 1;
    int sm_count = 0;
    int xyp_state = 0;   // tensor-map layout the driver accepted for k_sweep_xyp (0 / 1), -1: refused
    long xyp_used = 0;   // launches of k_sweep_xyp
    long xyu_used = 0;   // launches of k_sweep_xyu
    // top of the part (z + 1 of the highest active cell), reduced by k_build_code_v: the single-GPU z sweep solves only
    // the cells below it (launch_sweep_zt).  -1: unknown (cell-form code build); pending: the copy to h_ztop is in flight
    int *d_ztop = nullptr, *h_ztop = nullptr;
    int ztop = -1, xlo = 0, xhi = 0;   // xlo / xhi: first x plane with an active cell / last + 1 (valid when ztop >= 0)
    bool ztop_pending = false;
    long ztrim_used = 0;  // z sweeps launched on trimmed lines
    long xtrim_used = 0;  // x sweeps launched on the x extent of the part only
    int maskv_used = 0;  // bit 0 / 1 / 2: the last code build / code transposes / pack build ran in word form (adi_mask_core.h)
    // per-kernel timing (adi_profile_*): 5 events per step, read lazily
    std::vector<cudaEvent_t> prof_ev;
    long prof_steps = 0;

    // ---- Cartesian ----
    bool cart_bound = false;
    int nx = 0, ny = 0, nz = 0;
    double dx = 0.0;
    const uint8_t *d_mask = nullptr;
    adi::Pack pack[3];
    bool scalar_robin = false;
    double face_coeff[6] = {0, 0, 0, 0, 0, 0};
    uint8_t *code[3] = {nullptr, nullptr, nullptr};  // code[a] may alias code[0]
    uint8_t *code_buf[3] = {nullptr, nullptr, nullptr};
    size_t code_cells = 0;
    bool code_dirty = true;
    adi::TileList tiles[3];
    uint8_t *d_tflags = nullptr, *h_tflags = nullptr;
    size_t tflags_cap = 0;
    // transposed copies for the x / y sweeps (line axis fastest, padded to npadT[a]); option xy2
    uint8_t *codeT[2] = {nullptr, nullptr};
    size_t codeT_bytes[2] = {0, 0};
    int npadT[2] = {0, 0};
    // surface-only coefficient fields (k_check_sparse): re-examined after a pack or mask change
    bool sparse[3] = {false, false, false};
    bool sparse_dirty = true;
    long sparse_trust = 0;  // bit a: the caller vouches for pack a (built by adi_cart_build_packs for the bound
                            // mask and untouched since); cleared by every pack / mask call
    unsigned long long *d_viol = nullptr, *h_viol = nullptr;  // [3] each
    long operand_epoch = 0;  // counts mask / pack / halo (re)bindings: keys the caches of the multi-GPU z solve
    adi::DistState *dist = nullptr;
    // z-slab decomposition: mask planes of the adjacent slabs (borrowed), scratch for the ghosts
    int slab_rank = 0, slab_nranks = 1;
    const uint8_t *d_mask_lo = nullptr, *d_mask_hi = nullptr;
    double *d_ghost = nullptr;
    size_t ghost_lines = 0;
    int *d_maxk = nullptr;  // k_spike_pack: longest reach of a unit-ghost response
    // host-array convenience path
    double *stage[2] = {nullptr, nullptr};
    size_t stage_cells = 0;
    // page-locked staging of pageable host arrays (stage_h2d / stage_d2h)
    void *pin[2] = {nullptr, nullptr};
    cudaEvent_t pin_ev[2] = {nullptr, nullptr};
    // pipelined host-array path: two slots, each with its own in/out staging pair
    double *pipe[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    size_t pipe_cells[2] = {0, 0};

    // ---- cylindrical ----
    bool cyl_bound = false;
    int nr = 0, nphi = 0, cnz = 0, nz_pitch = 0;
    double dr = 0.0, dphi = 0.0, dz = 0.0;
    adi::CylTables *cyl = nullptr;
    uint8_t *stage_mask = nullptr;
    double *stage_src = nullptr;
    size_t stage_aux_cells = 0;

    // ---- output path (own streams and buffers; may run beside the stepping thread) ----
    adi::TextState *text = nullptr;
};
