/*
 * adi_b200.h -- C ABI of libadi_b200.so, the B200 (sm_100a) ADI heat-step engine.
 *
 * The reference (Matemusi/ADI_thermal_fields) has no FFI: its hot path is reached
 * through Python module functions.  Each entry point below names the reference
 * interface it stands behind (file:line relative to the reference tree); the
 * Python glue in adi_thermal_fields_b200/ binds them with ctypes and re-exports
 * the reference's own names.
 *
 * Conventions
 *   - every function returns 0 on success, a negative ADI_E* code otherwise;
 *     adi_last_error() returns a thread-local message for the last failure.
 *   - `d_*` pointers are DEVICE pointers borrowed for the duration of the call
 *     (or until replaced, for bound operands); `h_*` pointers are host pointers.
 *   - arrays are C-order: Cartesian (nx,ny,nz) and cylindrical (nr,nphi,nz) have
 *     z contiguous; fields are IEEE fp64, masks are 1 byte per cell (0 / non-0).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the default stream).
 *     Calls are asynchronous on that stream unless stated otherwise.
 *   - no CPU fallback exists anywhere behind this ABI.
 */
#ifndef ADI_B200_H
#define ADI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ADI_OK 0
#define ADI_EINVAL (-1)   /* bad argument / shape / unknown kind (ValueError in the reference) */
#define ADI_ECUDA (-2)    /* a CUDA runtime call or kernel launch failed */
#define ADI_ESTATE (-3)   /* call order violated (e.g. step before bind) */
#define ADI_ENOMEM (-4)

typedef struct adi_ctx adi_ctx;

/* ---- context / memory ------------------------------------------------------------- */
/* One context per process and GPU (SURVEY 8b "Threading").  Owns scratch buffers. */
int adi_ctx_create(int device, adi_ctx **out);
int adi_ctx_destroy(adi_ctx *ctx);
const char *adi_last_error(void);
const char *adi_version(void);
/* cp.cuda.Stream.null.synchronize() (quick_compare_neumann_robin_backend.py:184) */
int adi_sync(adi_ctx *ctx, void *stream);
/* Plain device memory for hosts without their own allocator (cp.full / cp.asarray / cp.asnumpy,
 * quick_compare_neumann_robin_backend.py:140-141,163-164). */
int adi_malloc(adi_ctx *ctx, size_t bytes, void **d_ptr);
int adi_free(adi_ctx *ctx, void *d_ptr);
int adi_h2d(adi_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, void *stream);
int adi_d2h(adi_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, void *stream);

/* ---- Cartesian path --------------------------------------------------------------- */
/* Grid3D(nx,ny,nz,dx,mask)  adi3d_gpu_coeff.py:6-12 / adi3d_numba_coeff.py:14-19 */
int adi_cart_bind(adi_ctx *ctx, int nx, int ny, int nz, double dx);
/* grid.mask (re)binding or in-place mutation (quick_compare_layer_birth_robin_v3.py:277,
 * waam_from_stl_v7_mm.py:494-495).  Borrowed device pointer; must stay valid until replaced.
 * Marks the per-cell neighbour code dirty (rebuilt lazily by the next step). */
int adi_cart_set_mask(adi_ctx *ctx, const uint8_t *d_mask);
/* AxisCoeffPack(coeff, dir_mask, dir_val, qflux)  adi3d_gpu_coeff.py:22-29, for axis 0/1/2.
 * NULL d_coeff / d_qflux mean all-zero; NULL d_dir_mask means no Dirichlet cell
 * (d_dir_val is then ignored).  Borrowed device pointers. */
int adi_cart_set_pack(adi_ctx *ctx, int axis, const double *d_coeff, const uint8_t *d_dir_mask,
                      const double *d_dir_val, const double *d_qflux);
/* Scalar Robin on the six faces ('x-','x+','y-','y+','z-','z+'): the coefficient
 * h*A/Ccell of adi3d_numba_coeff.py:93-99 is derived from the mask inside the sweep instead
 * of being read from a dense array.  face_coeff[f] = h_f*dx^2/(rho*cp*dx^3) as the reference
 * evaluates it.  Replaces the `coeff` operand of all three packs until set_pack is called. */
int adi_cart_set_robin_scalar(adi_ctx *ctx, const double face_coeff[6]);
/* adi_step_gpu_coeff(Tn, grid, mat, params, packs, Tinf)  adi3d_gpu_coeff.py:213-230
 * (same mathematics as adi_step_numba_coeff, adi3d_numba_coeff.py:290-302).
 * d_Tin is not modified; d_Tout must not alias it.  kappa = k/(rho*cp). */
int adi_cart_step(adi_ctx *ctx, const double *d_Tin, double *d_Tout, double dt, double theta,
                  double kappa, double Tinf, void *stream);
/* The same step for callers holding HOST arrays (the CPU module's calling convention,
 * adi3d_numba_coeff.py:290): copies h_Tin to the device, steps `nsteps` times, copies the
 * result back and synchronises.  Used for end-to-end timing. */
int adi_cart_step_host(adi_ctx *ctx, const double *h_Tin, double *h_Tout, int nsteps, double dt,
                       double theta, double kappa, double Tinf, void *stream);
/* Pipelined form of the above for callers that step many independent host fields (or stream
 * frames out while the next input streams in): asynchronous on `stream`, no synchronisation,
 * staging buffers of `slot` (0 or 1).  Two slots on two streams overlap the upload of one
 * step with the compute of the other and the download of the previous result (both PCIe
 * directions busy).  h_Tin / h_Tout should be pinned; the caller synchronises the stream
 * (adi_sync) before reading h_Tout or reusing the slot's host buffers.  The neighbour code
 * must be current (one adi_cart_step / adi_cart_step_host call after the last mask change)
 * before two streams are used concurrently. */
int adi_cart_step_host_async(adi_ctx *ctx, int slot, const double *h_Tin, double *h_Tout, double dt,
                             double theta, double kappa, double Tinf, void *stream);
/* precompute_coeff_packs_unified  adi3d_gpu_coeff.py:50-110 on the device.
 * For face f: h_kind[f] 0 = no Robin, 1 = scalar h_scalar[f], 2 = device field d_h_field[f];
 * likewise q_* for Neumann (q'' > 0 heats the solid).  Writes the six dense outputs
 * (any of which may be NULL to skip it).  Uses the mask bound with adi_cart_set_mask. */
int adi_cart_build_packs(adi_ctx *ctx, double rho, double cp, const int h_kind[6],
                         const double h_scalar[6], const double *const d_h_field[6],
                         const int q_kind[6], const double q_scalar[6],
                         const double *const d_q_field[6], double *d_coeff_x, double *d_coeff_y,
                         double *d_coeff_z, double *d_q_x, double *d_q_y, double *d_q_z,
                         void *stream);
/* exposed_mask(mask, face)  adi3d_gpu_coeff.py:31-48; face index 0..5 as above. */
int adi_cart_exposed_mask(adi_ctx *ctx, int face, uint8_t *d_out, void *stream);
/* ---- z-slab decomposition of the Cartesian grid across GPUs (BASELINE north_star; the
 * reference is single-process: this is the multi-GPU form of adi_step_gpu_coeff,
 * adi3d_gpu_coeff.py:213-230).  Rank r of R holds the planes [z0_r, z1_r) of every array; the
 * context is bound (adi_cart_bind) with the LOCAL nz.  x and y sweeps are local.  Per step the
 * host exchanges one T plane per side (explicit stage, adi3d_numba_coeff.py:274-288) and the
 * interface relations of the z sweep (adi3d_numba_coeff.py:205-237 solved as a partitioned
 * tridiagonal system), e.g. with NCCL:
 *     adi_cart_pack_zplanes(T)            -> send lo plane down / hi plane up
 *     adi_cart_step_xy(Tin, Tout, Tlo, Thi)
 *     adi_cart_zsweep_reduce(Tout, dyn, stat) -> all-gather dyn (2*nx*ny doubles per rank) and, when the
 *                                              matrix changed, stat (4*nx*ny)
 *     adi_cart_zsweep_finish(Tout, dyn_all, stat_all)
 * The local nz must be a multiple of 16 (32 when nz > 512). */
int adi_cart_set_slab(adi_ctx *ctx, int rank, int nranks);
/* Mask planes (nx*ny bytes) of the slab below / above; NULL at the domain boundary.  Borrowed.
 * Used by the neighbour code, adi_cart_build_packs and adi_cart_exposed_mask. */
int adi_cart_set_mask_halo(adi_ctx *ctx, const uint8_t *d_mask_lo, const uint8_t *d_mask_hi);
/* First / last z plane of a field (elem_bytes 8) or mask (1) into contiguous nx*ny buffers
 * (either output may be NULL). */
int adi_cart_pack_zplanes(adi_ctx *ctx, const void *d_field, int elem_bytes, void *d_lo_out,
                          void *d_hi_out, void *stream);
/* Explicit stage + x sweep + y sweep; d_Tlo / d_Thi: T planes received from the adjacent ranks. */
int adi_cart_step_xy(adi_ctx *ctx, const double *d_Tin, double *d_Tout, const double *d_Tlo,
                     const double *d_Thi, double dt, double theta, double kappa, double Tinf,
                     void *stream);
/* z sweep, pass 1: the interface relation of every local line segment,
 *   x_first = yf + vf*L + wf*R,  x_last = yl + vl*L + wl*R   (L, R: the adjacent ranks' boundary values),
 * as d_iface_dyn[2][nx*ny] = (yf, yl) and d_iface_stat[4][nx*ny] = (vf, wf, vl, wl).  The second
 * group depends on the matrix only (mask, packs, dt, theta): pass d_iface_stat = NULL to compute the
 * right-hand-side part alone while those stay the same.  d_T is not modified. */
int adi_cart_zsweep_reduce(adi_ctx *ctx, double *d_T, double *d_iface_dyn, double *d_iface_stat,
                           double dt, double theta, double kappa, double Tinf, void *stream);
/* z sweep, pass 2: d_dyn_all[nranks][2][nx*ny] and d_stat_all[nranks][4][nx*ny] gathered from all
 * ranks; solves the inter-rank system per line and finishes the local segments in place. */
int adi_cart_zsweep_finish(adi_ctx *ctx, double *d_T, const double *d_dyn_all, const double *d_stat_all,
                           double dt, double theta, double kappa, double Tinf, void *stream);
/* "Solve first" form of the same split, for steady stepping (mask, packs, dt, theta unchanged for many steps).
 * x = y + L*v + R*w per local line segment: y = the segment solved with both ghosts at zero, v / w = its
 * responses to a unit ghost at the lower / upper end (matrix only), L / R = the ghost values from the
 * gathered relations.  v and w decay geometrically away from their end (ratio ~ theta*gamma / (1 + 2 theta*gamma)),
 * so only the `kmax` cells next to each end are kept and corrected: the second full pass over the slab goes away.
 *   adi_cart_zsweep_spike   once per (mask, packs, dt, theta): computes the response of end 0 / 1 in the zeroed
 *                           scratch field (size of T), keeps compact[nx*ny][kmax] and K[nx*ny] (cells above
 *                           `threshold`, clipped to kmax); *h_maxK = longest reach found -- if it exceeds kmax the
 *                           caller stays with adi_cart_zsweep_reduce / _finish.  Synchronises.  Bit 30 of K[line] marks
 *                           a line whose response equals, bit for bit, that of the mid-grid line: adi_cart_zsweep_apply
 *                           then reads that one row for all of them (the count is K[line] & 0x3fffffff).
 *   adi_cart_zsweep_solve0  every step: solves T in place with zero ghosts, writes d_iface_dyn[2][nx*ny]
 *   adi_cart_zsweep_apply   every step, after the all-gather: ghosts from the relations, then the corrections */
int adi_cart_zsweep_spike(adi_ctx *ctx, double *d_scratch, int end, int kmax, double threshold, double *d_compact,
                          int *d_K, int *h_maxK, double dt, double theta, double kappa, void *stream);
int adi_cart_zsweep_solve0(adi_ctx *ctx, double *d_T, double *d_iface_dyn, double dt, double theta, double kappa,
                           double Tinf, void *stream);
int adi_cart_zsweep_apply(adi_ctx *ctx, double *d_T, const double *d_dyn_all, const double *d_stat_all,
                          const double *d_vC, const double *d_wC, const int *d_Kv, const int *d_Kw, int kmax,
                          void *stream);
/* Tuning / introspection: kernel variant selection (0 = default) and launch counter.  Options:
 *   "kt"    lanes along z per block of the strided sweeps (power of two)     "lt"  lines per block of the z sweeps
 *   "m"     chunk length 16 | 32                                             "wide" 1: 512-thread blocks for lines <= 512 cells
 *   "fuse"  1: explicit stage fused into the x sweep instead of its own pass
 *   "sparse_coeff" 1 (default): after every pack / mask change the dense coefficient fields of the x and y packs are
 *           examined (one pass, one small read-back); a field that is +0.0 on every active cell with both neighbours
 *           along its axis -- as everything precompute_coeff_packs_unified builds is -- is then read at exposed cells
 *           only.  Same bits as the dense reads; 0 switches the examination off
 *   "profile" 1: record per-kernel CUDA events (adi_profile_read)            "sync_check" 1: synchronise after every step
 *   "maskv" 1 (default): the kernels that run once per mask change (neighbour code, its transposed copies, the pack
 *           builder of adi_cart_build_packs) take their word-at-a-time forms where nz and the addresses allow
 *           (nz % 16 / 4 / 2 == 0); 0: one cell per thread.  Same bits either way
 *   "ztrim" 1 (default): a part under construction along z (waam_from_stl_v7_mm.py:487-550) -- the single-GPU z sweep
 *           solves only the cells below the highest active plane (void cells above it are identity rows), the x sweep
 *           only the x planes that hold an active cell
 *   read-only: "maskv_used" (bit 0 / 1 / 2: code / transposes / packs last ran in word form), "ztop" (highest active
 *           plane + 1, -1 unknown), "xlo" / "xhi" (first x plane with an active cell / last + 1), "ztrim_used" /
 *           "xtrim_used" (trimmed z / x sweeps so far) */
int adi_set_option(adi_ctx *ctx, const char *name, long value);
long adi_get_option(adi_ctx *ctx, const char *name);  /* value of an option; "sparse_active": bit a set when the
                                                          sweep along axis a reads its coefficient field at exposed
                                                          cells only (state after the last step); -1 = unknown name */
long adi_launch_count(adi_ctx *ctx);
/* Per-kernel device timing (the `[time]` prints of quick_compare_neumann_robin_backend.py:
 * 173-186 are the reference's only profiling hook).  After adi_set_option(ctx,"profile",1)
 * every adi_cart_step / adi_cyl_step records CUDA events around its kernels on the caller's
 * stream; adi_profile_read synchronises, returns the accumulated milliseconds per kernel
 * (explicit stage, x|r sweep, y|phi sweep, z sweep) and the number of steps since the last
 * adi_profile_reset. */
int adi_profile_reset(adi_ctx *ctx);
int adi_profile_read(adi_ctx *ctx, double ms[4], long *nsteps);

/* ---- multi-GPU: the z-slab step sequenced inside the library over NCCL (SURVEY 8b "adi_dist_init", 8e) ----
 * The reference is single-process; this is the multi-GPU form of adi_step_gpu_coeff (adi3d_gpu_coeff.py:213-230): one
 * process per GPU, rank r of R holds the z planes [z0_r, z1_r) of every array and binds its LOCAL grid
 * (adi_cart_bind with the local nz, a multiple of 16; 32 for local nz > 1024), mask and packs as on a single GPU.
 * NCCL is bound at run time (dlopen libnccl.so.2); nothing here is needed on a single GPU.
 *   adi_dist_unique_id   rank 0: a fresh 128-byte communicator id (ncclGetUniqueId); the host carries it to the
 *                        other ranks by its own means (MPI, a file, torch.distributed ...)
 *   adi_dist_init        every rank: joins the communicator (ncclCommInitRank); collective
 *   adi_dist_init_comm   alternative: adopt an ncclComm_t the host already owns (not destroyed by the library)
 *   adi_dist_set_option  "batches" (line batches of the overlapped z solve, default 4; "batch_min_lines": no batch smaller than
 *                        this, default 1048576), "spike_after" (steps with
 *                        unchanged operands before the solve-first z form replaces the two-pass form, default 2;
 *                        < 0 never), "spike_kmax" (reach of the ghost corrections in cells, default 32),
 *                        "spike_thr_log2" (responses below 2^value of the ghost are dropped, default -60),
 *                        "overlap_halo" (1: explicit stage beside the T-plane exchange, default 0)
 *   adi_dist_info        rank / size and how many steps ran in either z form */
int adi_dist_unique_id(void *id128);
int adi_dist_init(adi_ctx *ctx, const void *id128, int rank, int nranks);
int adi_dist_init_comm(adi_ctx *ctx, void *nccl_comm, int rank, int nranks);
int adi_dist_comm(adi_ctx *ctx, void **nccl_comm);   /* the ncclComm_t in use (e.g. to share it with a second context) */
int adi_dist_destroy(adi_ctx *ctx);
int adi_dist_set_option(adi_ctx *ctx, const char *name, long value);
int adi_dist_info(adi_ctx *ctx, int *rank, int *nranks, long *steps_two_pass, long *steps_solve_first);
/* Collective, after adi_cart_set_mask and whenever the mask changes (layer births, waam_from_stl_v7_mm.py:487-495):
 * the adjacent ranks' boundary mask planes are exchanged, so that faces on a slab boundary couple / are exposed
 * exactly as in the undivided grid (adi_cart_set_slab + adi_cart_set_mask_halo are applied inside). */
int adi_cart_slab_sync_mask(adi_ctx *ctx, void *stream);
/* Collective: one theta-step on the local slab.  Per step the ranks exchange one T plane per side (explicit stage,
 * adi3d_numba_coeff.py:274-288) and all-gather two doubles per z line and rank (partitioned z solve, :205-237), on
 * the library's own communication stream; the z solve runs in line batches so that the gather of one batch overlaps
 * the solve of the next.  Results equal adi_cart_step on the undivided grid to rounding. */
int adi_cart_slab_step(adi_ctx *ctx, const double *d_Tin, double *d_Tout, double dt, double theta, double kappa,
                       double Tinf, void *stream);

/* ---- voxel_bc_correction (producer of the per-face dense Robin fields) ------------------
 * STLBoundaryCorrector.compute_voxel_projected_areas  voxel_bc_correction.py:53-108: triangles
 * (d_tri[ntri][3][3]), unit normals (d_nrm[ntri][3]) and areas (d_area[ntri]) of the surface mesh are
 * subdivided and scattered into six per-face projected-area fields d_proj[f] (face order
 * 'x-','x+','y-','y+','z-','z+'; the caller zero-fills them).  fp64 atomics: the sum into one voxel is
 * order-dependent at rounding level. */
int adi_voxel_project(adi_ctx *ctx, const double *d_tri, const double *d_nrm, const double *d_area,
                      int ntri, const double origin[3], double dx, int max_subdiv, double area_eps,
                      const uint8_t *d_mask, int nx, int ny, int nz, double *const d_proj[6],
                      void *stream);
/* STLBoundaryCorrector.build_corrected_fields  voxel_bc_correction.py:110-168: for every face with
 * has[f] != 0, robin[f] = base_h[f] * proj[f] / dx^2 and scale[f] = proj[f] / dx^2; with `fallback`,
 * exposed faces without projected area take base_h[f] / 1.  d_scale (or single entries) may be NULL. */
int adi_voxel_correct(adi_ctx *ctx, const uint8_t *d_mask, int nx, int ny, int nz, double dx,
                      const double *const d_proj[6], const int has[6], const double base_h[6],
                      int fallback, double *const d_robin[6], double *const d_scale[6], void *stream);

/* ---- cylindrical path --------------------------------------------------------------- */
/* GridCyl(nr,nphi,nz,dr,dphi,dz,R)  adi3d_cyl_phi_v3.py:33-43.  nz_pitch >= nz is the
 * allocated z extent of the fields (layer births grow nz without reallocating,
 * quick_compare_layer_birth_robin_cyl_v3.py:195-204). */
int adi_cyl_bind(adi_ctx *ctx, int nr, int nphi, int nz, int nz_pitch, double dr, double dphi,
                 double dz);
#define ADI_Z_NEUMANN0 0
#define ADI_Z_DIRICHLET 1
#define ADI_Z_ROBIN 2
typedef struct adi_cyl_params {
    double dt;                /* Params.dt            adi3d_cyl_phi_v3.py:52-54 */
    double rho, cp, k;        /* Material             :45-50 */
    double h_r, Tinf_r;       /* RobinR               :56-58 */
    int kind_bot, kind_top;   /* ZBC kinds            :60-68 */
    double h_bot, h_top, Tinf_bot, Tinf_top, T_bot, T_top;
    double T_void, T_inner;   /* adi_step_masked clamps, quick_spiral_deposition_gif_v5.py:51-68 */
} adi_cyl_params;
/* adi_step(Tn, grid, mat, prm, robin_r, zbc, S)  adi3d_cyl_phi_v3.py:332-350 (scheme "be"),
 * and, when d_active != NULL, adi_step_masked  quick_spiral_deposition_gif_v5.py:31-70.
 * d_S may be NULL (no source). */
int adi_cyl_step(adi_ctx *ctx, const double *d_Tin, double *d_Tout, const adi_cyl_params *p,
                 const uint8_t *d_active, const double *d_S, void *stream);
int adi_cyl_step_host(adi_ctx *ctx, const double *h_Tin, double *h_Tout, int nsteps,
                      const adi_cyl_params *p, const uint8_t *h_active, const double *h_S,
                      void *stream);
/* z-slab decomposition of the cylindrical grid (multi-GPU form of adi_step, adi3d_cyl_phi_v3.py:332-350):
 * rank r of R holds the z planes [z0_r, z1_r) of every array and is bound (adi_cyl_bind) with its LOCAL
 * nz; nz_per_rank lists the local nz of all ranks.  The r and phi solves are rank-local; the z solve
 * (:348-350) is partitioned: pass 1 gives, per line, the right-hand-side part d_y[2][nr*nphi] = (yf, yl) of
 * the segment's interface relation (its matrix part is line-independent and tabulated on the host), the
 * ranks all-gather them into d_y_all[R][2][nr*nphi], pass 2 maps them to the two ghost values of every
 * line and finishes the segment.  The boundary rows of ZBC apply on the first / last rank only. */
int adi_cyl_set_slab(adi_ctx *ctx, int rank, int nranks, const int *nz_per_rank);
int adi_cyl_step_rphi(adi_ctx *ctx, const double *d_Tin, double *d_Tout, const adi_cyl_params *p,
                      const uint8_t *d_active, const double *d_S, void *stream);
int adi_cyl_zsweep_reduce(adi_ctx *ctx, double *d_T, const adi_cyl_params *p, double *d_y, void *stream);
int adi_cyl_zsweep_finish(adi_ctx *ctx, double *d_T, const adi_cyl_params *p, const double *d_y_all,
                          const uint8_t *d_active, void *stream);

/* ---- output path ------------------------------------------------------------
 * The reference writes its fields with per-value Python formatting:
 *   fmt 0  vtk_writer.py:4-30  write_vtk_structured_points: "{float(v):.6e}", values in Fortran
 *          order of the C-order array (x fastest), nine per line;
 *   fmt 1  waam_from_stl_v7_mm.py:186-215 write_vtk_structured_points: "{float(T[i,j,k]):.6g}",
 *          one line per (k, j) row of nx values.
 * Here the device produces exactly those bytes (correctly rounded, ties to even, "nan"/"inf"
 * spelled as Python does).  dtype: 0 = fp64, 1 = fp32, 2 = 1-byte mask (0/1).
 * The output path keeps its own streams and buffers: these calls may be issued from a second
 * host thread while the first one keeps stepping. */
size_t adi_text_capacity(size_t nvalues);  /* bytes that always hold the text of nvalues values */
/* Text of planes [k0, k0+kc) of the field into d_text (device, 16-byte aligned); *nbytes = its
 * size.  Returns when the text is complete. */
int adi_text_format(adi_ctx *ctx, const void *d_field, int dtype, int nx, int ny, int nz, int k0, int kc,
                    int fmt, char *d_text, size_t capacity, unsigned long long *nbytes, void *stream);
/* Opens `path` (truncate, or append when `append` != 0), writes `prefix` (the caller's header
 * lines) and then the value lines of the whole field, formatted on the device in plane chunks
 * and copied through pinned double buffers while the previous chunk is being written.  The
 * field is read after everything queued on `stream` so far.  Returns when the file is closed. */
int adi_text_write(adi_ctx *ctx, const char *path, int append, const void *prefix, size_t prefix_len,
                   const void *d_field, int dtype, int nx, int ny, int nz, int fmt,
                   unsigned long long *bytes_written, void *stream);
/* Probe lines / slices / boxes (T[i0,j0,:], T[:,j0,:] ... in the drivers,
 * quick_compare_neumann_robin_backend.py:147-150, waam_from_stl_v7_mm.py:497-513) without stalling
 * the stepping stream: adi_probe_record packs the box [lo, hi) of a C-order (nx,ny,nz) array of
 * elem_bytes (8/4/1) items on `stream` and hands the PCIe copy to a copy stream; adi_probe_fetch
 * copies the landed record out (wait != 0 blocks until it is there; with wait == 0 it returns 1
 * while the copy is still in flight).  A slot holds one record until it is fetched. */
int adi_probe_open(adi_ctx *ctx, int nslots, size_t slot_bytes);
int adi_probe_record(adi_ctx *ctx, int slot, const void *d_field, int elem_bytes, int nx, int ny, int nz,
                     const int lo[3], const int hi[3], void *stream);
int adi_probe_fetch(adi_ctx *ctx, int slot, void *h_dst, size_t capacity, size_t *nbytes, int wait);

#ifdef __cplusplus
}
#endif
#endif /* ADI_B200_H */
