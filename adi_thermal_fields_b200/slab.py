# -*- coding: utf-8 -*-
"""z-slab decomposition of the Cartesian ADI step across GPUs (BASELINE north_star).

The reference is single-process; this is the multi-GPU form of `adi_step_gpu_coeff`
(adi3d_gpu_coeff.py:213-230): rank r of R owns the z planes [z0_r, z1_r) of every array.

  * x and y sweeps are rank-local (lines never cross slabs);
  * the explicit stage (adi3d_numba_coeff.py:274-288,298) needs one T plane from each adjacent
    rank per step, the neighbour code one MASK plane per side whenever the mask changes;
  * the z sweep (adi3d_numba_coeff.py:205-237) is a partitioned tridiagonal solve: every rank
    reduces its segment of each line to an interface relation (pass 1), the relations are
    all-gathered (2 doubles per line and rank per step, plus 4 matrix-only doubles whenever mask,
    packs, dt or theta change), every rank solves the 2R-unknown inter-rank system per line and
    finishes its segment (pass 2).

One process per GPU; `torch.distributed` (NCCL over NVLink) is the plumbing for the two
exchanges, the arithmetic runs in libadi_b200.so.  `LocalComm` runs R virtual ranks as threads of
one process on one device (used by the single-GPU tests and to measure the cost of the split
itself); `backend=` is a test seam (tests/ plug in the host emulation of the kernels for the
gloo tests on CPU) -- the default backend is the CUDA library and raises without it.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np
import torch

from . import _capi

FACES = ("x-", "x+", "y-", "y+", "z-", "z+")
USE_LIBRARY_SEQUENCING = True   # NCCL ranks: adi_cart_slab_step (False: the call-by-call sequencing of this module)


def split_z(nz, world, multiple=1):
    """Balanced slab extents: list of (z0, z1).  With `multiple` every extent but the last is a multiple of
    it (the Cartesian z-sweep kernels of the slab path need local nz % 16 == 0, 32 for local nz > 1024; the last
    rank takes the remainder and must then satisfy it too -- i.e. nz itself must be a multiple)."""
    if multiple > 1:
        units, rem = divmod(nz, multiple)
        q, r = divmod(units, world)
        out, z = [], 0
        for i in range(world):
            n = (q + (1 if i < r else 0)) * multiple + (rem if i == world - 1 else 0)
            out.append((z, z + n))
            z += n
        return out
    q, r = divmod(nz, world)
    out, z = [], 0
    for i in range(world):
        n = q + (1 if i < r else 0)
        out.append((z, z + n))
        z += n
    return out


# --------------------------------------------------------------------------------------------
# communicators
# --------------------------------------------------------------------------------------------
class TorchDistComm:
    """Ranks = processes of a torch.distributed group (NCCL for CUDA tensors, gloo for CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)

    def exchange_planes(self, lo, hi, recv_lo, recv_hi):
        """Send `hi` up / `lo` down; receive the plane below into recv_lo, above into recv_hi."""
        dist = self.dist
        ops = []
        if self.rank + 1 < self.world:
            ops.append(dist.P2POp(dist.isend, hi, self._peer(self.rank + 1), self.group))
            ops.append(dist.P2POp(dist.irecv, recv_hi, self._peer(self.rank + 1), self.group))
        if self.rank > 0:
            ops.append(dist.P2POp(dist.isend, lo, self._peer(self.rank - 1), self.group))
            ops.append(dist.P2POp(dist.irecv, recv_lo, self._peer(self.rank - 1), self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    def _peer(self, r):
        return r if self.group is None else self.dist.get_global_rank(self.group, r)

    def all_gather(self, out, inp):
        try:
            self.dist.all_gather_into_tensor(out, inp, group=self.group)
        except (RuntimeError, NotImplementedError):
            parts = list(out.view(self.world, -1).unbind(0))
            self.dist.all_gather(parts, inp.view(-1), group=self.group)


class LocalComm:
    """R virtual ranks = R threads of this process sharing one device (and its current stream, so
    every hand-over is stream-ordered).  comm = LocalComm(R); comm.view(r) is rank r's handle."""

    def __init__(self, world):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [dict() for _ in range(world)]

    def view(self, rank):
        return _LocalView(self, rank)

    def run(self, fn):
        """Run fn(rank_view) on every virtual rank; returns the list of results."""
        res, err = [None] * self.world, []
        dev = torch.cuda.current_device() if torch.cuda.is_available() else None

        def go(r):
            try:
                if dev is not None:
                    torch.cuda.set_device(dev)
                res[r] = fn(self.view(r))
            except BaseException as e:  # noqa: BLE001
                err.append(e)
                self.barrier.abort()

        th = [threading.Thread(target=go, args=(r,)) for r in range(self.world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if err:
            raise err[0]
        return res


class _LocalView:
    def __init__(self, comm, rank):
        self.c, self.rank, self.world = comm, rank, comm.world

    def exchange_planes(self, lo, hi, recv_lo, recv_hi):
        s = self.c.slots
        s[self.rank]["lo"], s[self.rank]["hi"] = lo, hi
        self.c.barrier.wait()
        if self.rank > 0:
            recv_lo.copy_(s[self.rank - 1]["hi"])
        if self.rank + 1 < self.world:
            recv_hi.copy_(s[self.rank + 1]["lo"])
        self.c.barrier.wait()

    def all_gather(self, out, inp):
        s = self.c.slots
        s[self.rank]["ag"] = inp
        self.c.barrier.wait()
        o = out.view(self.world, -1)
        for r in range(self.world):
            o[r].copy_(s[r]["ag"].view(-1))
        self.c.barrier.wait()


# --------------------------------------------------------------------------------------------
# CUDA backend (the product path)
# --------------------------------------------------------------------------------------------
class CudaBackend:
    """One engine context per (virtual) rank; all arithmetic through the C ABI."""

    def __init__(self):
        if not torch.cuda.is_available():
            raise RuntimeError("adi_thermal_fields_b200.slab: no CUDA device (there is no CPU fallback)")
        self.L = _capi.load()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        h = C.c_void_p()
        _capi.check(self.L.adi_ctx_create(self.dev.index, C.byref(h)), "adi_ctx_create")
        self.ctx = h
        self.hold = None
        self.pack_key = None
        self.dist = False     # True: the library sequences the slab step over its own NCCL communicator

    def _st(self):
        return torch.cuda.current_stream().cuda_stream

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.dev)

    def asarray(self, x, dtype):
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        return t.to(device=self.dev, dtype=dtype).contiguous()

    def bind(self, nx, ny, nz, dx, mask, rank, world):
        L = self.L
        _capi.check(L.adi_cart_bind(self.ctx, nx, ny, nz, dx), "adi_cart_bind")
        _capi.check(L.adi_cart_set_slab(self.ctx, rank, world), "adi_cart_set_slab")
        _capi.check(L.adi_cart_set_mask(self.ctx, mask.data_ptr()), "adi_cart_set_mask")
        self.pack_key = None

    def pack_planes(self, field, lo, hi):
        _capi.check(self.L.adi_cart_pack_zplanes(self.ctx, field.data_ptr(), field.element_size(),
                                                 lo.data_ptr(), hi.data_ptr(), self._st()), "adi_cart_pack_zplanes")

    def set_mask_halo(self, lo, hi):
        _capi.check(self.L.adi_cart_set_mask_halo(self.ctx, None if lo is None else lo.data_ptr(),
                                                  None if hi is None else hi.data_ptr()), "adi_cart_set_mask_halo")

    def mark_mask_changed(self, mask):
        _capi.check(self.L.adi_cart_set_mask(self.ctx, mask.data_ptr()), "adi_cart_set_mask")

    def build_packs(self, rho, cp, hk, hs, hf, qk, qs, qf, shape):
        """-> (coeff[3], q[3]) device arrays or None (k_build_packs, halo-aware exposed faces)."""
        dense_h = any(k == 2 for k in hk)
        have_q = any(k != 0 for k in qk)
        if not (dense_h or have_q):
            return [None] * 3, [None] * 3
        coeffs = [self.empty(shape, torch.float64) if dense_h else None for _ in range(3)]
        qs_out = [self.empty(shape, torch.float64) if (qk[2 * a] or qk[2 * a + 1]) else None for a in range(3)]
        vp = C.c_void_p
        hfp = (vp * 6)(*[None if a is None else a.data_ptr() for a in hf])
        qfp = (vp * 6)(*[None if a is None else a.data_ptr() for a in qf])
        ptr = [None if a is None else a.data_ptr() for a in coeffs + qs_out]
        _capi.check(self.L.adi_cart_build_packs(self.ctx, rho, cp, (C.c_int * 6)(*hk), (C.c_double * 6)(*hs), hfp,
                                                (C.c_int * 6)(*qk), (C.c_double * 6)(*qs), qfp, *ptr, self._st()),
                    "adi_cart_build_packs")
        return coeffs, qs_out

    def set_packs(self, packs, face_coeff):
        """packs: 3 x (coeff|None, dir_mask|None, dir_val|None, q|None) device arrays."""
        # re-bound only when an array was replaced or written (tensor version): every adi_cart_set_pack makes
        # the engine re-examine the coefficient fields (option sparse_coeff), one pass and one read-back
        key = (tuple(None if x is None else (x.data_ptr(), x._version) for p in packs for x in p),
               None if face_coeff is None else tuple(face_coeff))
        if key == self.pack_key:
            return
        self.hold = packs            # keeps the keyed tensors alive
        for a, (c, dm, dv, q) in enumerate(packs):
            p = [None if x is None else x.data_ptr() for x in (c, dm, dv, q)]
            _capi.check(self.L.adi_cart_set_pack(self.ctx, a, *p), "adi_cart_set_pack")
        if face_coeff is not None:
            _capi.check(self.L.adi_cart_set_robin_scalar(self.ctx, (C.c_double * 6)(*face_coeff)),
                        "adi_cart_set_robin_scalar")
        self.pack_key = key

    def step_xy(self, Tin, Tout, Tlo, Thi, dt, theta, kappa, Tinf):
        _capi.check(self.L.adi_cart_step_xy(self.ctx, Tin.data_ptr(), Tout.data_ptr(),
                                            None if Tlo is None else Tlo.data_ptr(),
                                            None if Thi is None else Thi.data_ptr(), dt, theta, kappa, Tinf,
                                            self._st()), "adi_cart_step_xy")

    def step_plain(self, Tin, Tout, dt, theta, kappa, Tinf):
        _capi.check(self.L.adi_cart_step(self.ctx, Tin.data_ptr(), Tout.data_ptr(), dt, theta, kappa, Tinf,
                                         self._st()), "adi_cart_step")

    def zsweep_reduce(self, T, dyn, stat, dt, theta, kappa, Tinf):
        _capi.check(self.L.adi_cart_zsweep_reduce(self.ctx, T.data_ptr(), dyn.data_ptr(),
                                                  None if stat is None else stat.data_ptr(), dt, theta, kappa, Tinf,
                                                  self._st()), "adi_cart_zsweep_reduce")

    def zsweep_finish(self, T, dyn_all, stat_all, dt, theta, kappa, Tinf):
        _capi.check(self.L.adi_cart_zsweep_finish(self.ctx, T.data_ptr(), dyn_all.data_ptr(), stat_all.data_ptr(), dt,
                                                  theta, kappa, Tinf, self._st()), "adi_cart_zsweep_finish")

    # "solve first" z pass (adi_cart_zsweep_spike / _solve0 / _apply)
    def zsweep_spikes(self, shape, kmax, threshold, dt, theta, kappa):
        """-> (vC, wC, Kv, Kw) compact unit-ghost responses of the two ends, or None when a response
        reaches further than kmax cells (slow decay: large theta*gamma)."""
        nl = shape[0] * shape[1]
        scratch = self.empty(shape, torch.float64)
        out = []
        for end in (0, 1):
            comp = self.empty((nl, kmax), torch.float64)
            K = self.empty((nl,), torch.int32)
            mk = C.c_int(0)
            _capi.check(self.L.adi_cart_zsweep_spike(self.ctx, scratch.data_ptr(), end, kmax, threshold, comp.data_ptr(),
                                                     K.data_ptr(), C.byref(mk), dt, theta, kappa, self._st()),
                        "adi_cart_zsweep_spike")
            if mk.value > kmax:
                return None
            out += [comp, K]
        return out[0], out[2], out[1], out[3]

    def zsweep_solve0(self, T, dyn, dt, theta, kappa, Tinf):
        _capi.check(self.L.adi_cart_zsweep_solve0(self.ctx, T.data_ptr(), dyn.data_ptr(), dt, theta, kappa, Tinf,
                                                  self._st()), "adi_cart_zsweep_solve0")

    def zsweep_apply(self, T, dyn_all, stat_all, spikes, kmax):
        vC, wC, Kv, Kw = spikes
        _capi.check(self.L.adi_cart_zsweep_apply(self.ctx, T.data_ptr(), dyn_all.data_ptr(), stat_all.data_ptr(),
                                                 vC.data_ptr(), wC.data_ptr(), Kv.data_ptr(), Kw.data_ptr(), kmax,
                                                 self._st()), "adi_cart_zsweep_apply")

    def launch_count(self):
        return int(self.L.adi_launch_count(self.ctx))

    def set_option(self, name, value):
        if str(name) in ("batches", "batch_min_lines", "spike_after", "spike_kmax", "overlap_halo", "spike_thr_log2"):
            _capi.check(self.L.adi_dist_set_option(self.ctx, str(name).encode(), int(value)), "adi_dist_set_option")
            return
        _capi.check(self.L.adi_set_option(self.ctx, str(name).encode(), int(value)), "adi_set_option")

    # ---- sequencing inside the library over its own NCCL communicator (csrc/adi_dist.cu) ----
    _owners = {}     # (device, rank, world) -> context that owns the process's communicator

    def dist_init(self, comm):
        """Join the library's NCCL communicator for `comm`'s ranks (created once per process and shared by later
        contexts).  Collective the first time.  The 128-byte id travels over torch.distributed."""
        key = (self.dev.index, comm.rank, comm.world)
        owner = CudaBackend._owners.get(key)
        if owner is None:
            buf = (C.c_ubyte * 128)()
            if comm.rank == 0:
                _capi.check(self.L.adi_dist_unique_id(buf), "adi_dist_unique_id")
            t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=self.dev)
            comm.dist.broadcast(t, src=comm._peer(0), group=comm.group)
            raw = bytes(t.cpu().tolist())
            _capi.check(self.L.adi_dist_init(self.ctx, raw, comm.rank, comm.world), "adi_dist_init")
            CudaBackend._owners[key] = self.ctx
            self._keep_owner = None
        else:
            h = C.c_void_p()
            _capi.check(self.L.adi_dist_comm(owner, C.byref(h)), "adi_dist_comm")
            _capi.check(self.L.adi_dist_init_comm(self.ctx, h, comm.rank, comm.world), "adi_dist_init_comm")
        self.dist = True

    def slab_sync_mask(self):
        _capi.check(self.L.adi_cart_slab_sync_mask(self.ctx, self._st()), "adi_cart_slab_sync_mask")

    def slab_step(self, Tin, Tout, dt, theta, kappa, Tinf):
        _capi.check(self.L.adi_cart_slab_step(self.ctx, Tin.data_ptr(), Tout.data_ptr(), dt, theta, kappa, Tinf,
                                              self._st()), "adi_cart_slab_step")

    def dist_info(self):
        r, n, a, b = C.c_int(), C.c_int(), C.c_long(), C.c_long()
        _capi.check(self.L.adi_dist_info(self.ctx, C.byref(r), C.byref(n), C.byref(a), C.byref(b)), "adi_dist_info")
        return dict(rank=r.value, nranks=n.value, steps_two_pass=a.value, steps_solve_first=b.value)

    def profile(self, on):
        self.L.adi_set_option(self.ctx, b"profile", 1 if on else 0)
        self.L.adi_profile_reset(self.ctx)

    def profile_read(self):
        ms = (C.c_double * 4)()
        n = C.c_long()
        self.L.adi_profile_read(self.ctx, ms, C.byref(n))
        return [ms[i] for i in range(4)], n.value

    def close(self):
        if self.ctx is not None and self.ctx not in CudaBackend._owners.values():
            self.L.adi_ctx_destroy(self.ctx)
            self.ctx = None


# --------------------------------------------------------------------------------------------
# the slab stepper
# --------------------------------------------------------------------------------------------
class SlabGrid3D:
    """Local slab of a Cartesian grid: Grid3D(nx,ny,nz,dx,mask) (adi3d_gpu_coeff.py:6-12) with
    nz / mask the LOCAL extent, plus the communicator that links it to the adjacent slabs."""

    def __init__(self, nx, ny, nz_local, dx, mask_local, comm, backend=None):
        self.nx, self.ny, self.nz, self.dx = int(nx), int(ny), int(nz_local), float(dx)
        self.comm = comm
        self.rank, self.world = comm.rank, comm.world
        self.be = backend if backend is not None else CudaBackend()
        be = self.be
        self.mask = be.asarray(mask_local, torch.bool)
        assert tuple(self.mask.shape) == (self.nx, self.ny, self.nz)
        self.mask_version = 0
        self._stat_key = None   # (dt, theta, kappa, packs, mask version) the gathered matrix part belongs to
        # steady stepping: after `spike_after` steps with the same key the z sweep switches to the solve-first
        # form (one pass + corrections next to the slab faces); None = not built yet, False = responses too long
        self.spike_after, self.spike_kmax, self.spike_threshold = 2, 32, 2.0 ** -60
        self._spikes, self._stat_uses = None, 0
        be.bind(self.nx, self.ny, self.nz, self.dx, self.mask, self.rank, self.world)
        # NCCL ranks (one process per GPU): the whole step is sequenced inside the library (csrc/adi_dist.cu: its own
        # communicator and communication stream, batched / overlapped z solve).  LocalComm (virtual ranks on one
        # device) and the CPU test backend keep the sequencing below, which is the same algorithm call by call.
        if (USE_LIBRARY_SEQUENCING and backend is None and isinstance(comm, TorchDistComm) and self.world > 1
                and comm.dist.get_backend(comm.group) == "nccl"):
            be.dist_init(comm)
        else:
            pl = (self.nx, self.ny)
            self._m_lo, self._m_hi = be.empty(pl, torch.bool), be.empty(pl, torch.bool)      # to send
            self.mask_lo, self.mask_hi = be.empty(pl, torch.bool), be.empty(pl, torch.bool)  # received
            self._t_lo, self._t_hi = be.empty(pl, torch.float64), be.empty(pl, torch.float64)
            self.T_lo, self.T_hi = be.empty(pl, torch.float64), be.empty(pl, torch.float64)
            nl = self.nx * self.ny
            self.iface_dyn, self.iface_stat = be.empty((2, nl), torch.float64), be.empty((4, nl), torch.float64)
            self.dyn_all = be.empty((self.world, 2, nl), torch.float64)
            self.stat_all = be.empty((self.world, 4, nl), torch.float64)
        self.sync_mask()

    def z_form(self):
        """'solve-first' / 'two-pass' / 'single GPU': the form the last steps' z sweep took."""
        if self.world == 1:
            return "single GPU"
        if getattr(self.be, "dist", False):
            i = self.be.dist_info()
            return "solve-first" if i["steps_solve_first"] > 0 else "two-pass"
        return "solve-first" if self._spikes else "two-pass"

    def sync_mask(self):
        """Call after the local mask changed (layer births): refreshes the adjacent ranks' view of
        it.  Collective: every rank of the group must call it."""
        be = self.be
        self.mask_version += 1
        be.mark_mask_changed(self.mask)
        if getattr(be, "dist", False):
            be.slab_sync_mask()
            self._mask_synced = (self.mask.data_ptr(), self.mask._version)
            return
        be.pack_planes(self.mask, self._m_lo, self._m_hi)
        self.comm.exchange_planes(self._m_lo, self._m_hi, self.mask_lo, self.mask_hi)
        be.set_mask_halo(self.mask_lo if self.rank > 0 else None,
                         self.mask_hi if self.rank + 1 < self.world else None)
        self._mask_synced = (self.mask.data_ptr(), self.mask._version)


class SlabPacks:
    def __init__(self, packs, face_coeff):
        self.packs, self.face_coeff = packs, face_coeff


def precompute_coeff_packs_unified(grid, mat, dir_mask=None, dir_value=None, neumann=None, robin_h=None,
                                   robin_Tinf=None):
    """precompute_coeff_packs_unified (adi3d_gpu_coeff.py:50-110) for a slab: all field arguments
    are the LOCAL slabs of the global fields; faces at slab boundaries are exposed only where
    the adjacent rank's cell is void."""
    be = grid.be
    shape = (grid.nx, grid.ny, grid.nz)

    def classify(v):
        if v is None:
            return 0, 0.0, None
        if np.isscalar(v):
            return 1, float(v), None
        a = be.asarray(v, torch.float64)
        if tuple(a.shape) != shape:
            raise ValueError("field shape does not match the slab")
        return 2, 0.0, a

    hk, hs, hf = [0] * 6, [0.0] * 6, [None] * 6
    if robin_h is not None:
        for i, f in enumerate(FACES):
            v = robin_h.get(f, 0.0) if isinstance(robin_h, dict) else robin_h
            hk[i], hs[i], hf[i] = classify(v)
    qk, qs, qf = [0] * 6, [0.0] * 6, [None] * 6
    if neumann is not None:
        for f, v in neumann.items():
            if f not in FACES:
                raise ValueError("bad face")
            i = FACES.index(f)
            qk[i], qs[i], qf[i] = classify(v)
    dm = None if dir_mask is None else be.asarray(dir_mask, torch.bool)
    if dm is not None and not bool(dm.any()):
        dm = None
    dv = None
    if dm is not None:
        if dir_value is None:
            dv = torch.zeros_like(dm, dtype=torch.float64)
        elif np.isscalar(dir_value):
            dv = torch.full_like(dm, float(dir_value), dtype=torch.float64)
        else:
            dv = be.asarray(dir_value, torch.float64)
    coeffs, qouts = be.build_packs(mat.rho, mat.cp, hk, hs, hf, qk, qs, qf, shape)
    A = grid.dx * grid.dx
    Ccell = mat.rho * mat.cp * grid.dx ** 3
    face_coeff = None
    if coeffs[0] is None:
        face_coeff = [(hs[i] * A / Ccell) if hk[i] == 1 else 0.0 for i in range(6)]
    return SlabPacks([(coeffs[a], dm, dv, qouts[a]) for a in range(3)], face_coeff)


def adi_step_gpu_coeff(Tn, grid, mat, params, packs, Tinf=0.0, out=None):
    """One ADI theta-step of the slab (collective over grid.comm).  Returns a new local array."""
    be, comm = grid.be, grid.comm
    kappa = mat.k / (mat.rho * mat.cp)
    dt, theta = float(params.dt), float(params.theta)
    T = Tn
    if out is None:
        out = be.empty(tuple(T.shape), torch.float64)
    be.set_packs(packs.packs, packs.face_coeff)
    if grid.world == 1 and hasattr(be, "step_plain"):   # a single slab is the plain step
        be.step_plain(T, out, dt, theta, kappa, float(Tinf))
        return out
    if getattr(be, "dist", False):   # sequenced inside the library (adi_cart_slab_step)
        mk = (grid.mask.data_ptr(), grid.mask._version)
        if mk != grid._mask_synced:
            raise RuntimeError("slab.adi_step_gpu_coeff: grid.mask was edited (or rebound) without sync_mask(); the adjacent "
                               "ranks' view of it and the neighbour code are stale")
        be.slab_step(T, out, dt, theta, kappa, float(Tinf))
        return out
    lo_ok, hi_ok = grid.rank > 0, grid.rank + 1 < grid.world
    if theta != 1.0:   # the explicit stage is the only consumer of the T halo (beta = 0 at theta = 1)
        be.pack_planes(T, grid._t_lo, grid._t_hi)
        comm.exchange_planes(grid._t_lo, grid._t_hi, grid.T_lo, grid.T_hi)
    be.step_xy(T, out, grid.T_lo if lo_ok else None, grid.T_hi if hi_ok else None, dt, theta, kappa, float(Tinf))
    # the matrix part of the interface relations (4 of the 6 numbers per line) only changes with
    # mask, packs, dt or theta: it is computed and gathered once and reused while those stay the same
    # ... keyed on the operands' identity AND content version (set_packs' key: data_ptr + tensor version of every
    # bound array) and on the mask tensor itself, so that in-place edits of coeff / dir_mask / mask are seen
    mk = (grid.mask.data_ptr(), grid.mask._version)
    if mk != grid._mask_synced:
        raise RuntimeError("slab.adi_step_gpu_coeff: grid.mask was edited (or rebound) without sync_mask(); the adjacent "
                           "ranks' view of it and the neighbour code are stale")
    key = (dt, theta, kappa, id(packs), getattr(be, "pack_key", None), grid.mask_version, mk)
    fresh = key != grid._stat_key
    if fresh:
        grid._spikes, grid._stat_uses = None, 0
    else:
        grid._stat_uses += 1
        if (grid._spikes is None and grid._stat_uses >= grid.spike_after and hasattr(be, "zsweep_spikes")):
            # a purely local decision: both forms hand the same relation to the other ranks
            kmax = min(grid.spike_kmax, grid.nz)
            sp = be.zsweep_spikes((grid.nx, grid.ny, grid.nz), kmax, grid.spike_threshold, dt, theta, kappa)
            grid._spikes = (sp, kmax) if sp is not None else False
    if grid._spikes:
        be.zsweep_solve0(out, grid.iface_dyn, dt, theta, kappa, float(Tinf))
        comm.all_gather(grid.dyn_all, grid.iface_dyn)
        be.zsweep_apply(out, grid.dyn_all, grid.stat_all, *grid._spikes)
        return out
    be.zsweep_reduce(out, grid.iface_dyn, grid.iface_stat if fresh else None, dt, theta, kappa, float(Tinf))
    comm.all_gather(grid.dyn_all, grid.iface_dyn)
    if fresh:
        comm.all_gather(grid.stat_all, grid.iface_stat)
        grid._stat_key = key
        grid._stat_packs = packs   # keeps id(packs) from being reused by another object
    be.zsweep_finish(out, grid.dyn_all, grid.stat_all, dt, theta, kappa, float(Tinf))
    return out


# --------------------------------------------------------------------------------------------
# cylindrical grid: z-slab decomposition of adi3d_cyl_phi_v3.adi_step (scheme "be")
# --------------------------------------------------------------------------------------------
class SlabGridCyl:
    """Local slab of a cylindrical grid: GridCyl(nr,nphi,nz,dr,dphi,dz,R) (adi3d_cyl_phi_v3.py:33-43) with nz
    the LOCAL number of z planes; `nz_per_rank` lists the local nz of every rank of `comm`.
    The r and phi solves are rank-local; only the z solve exchanges data (one all-gather of two doubles per
    line and rank per step -- the matrix part of the interface relations is line-independent and lives in
    host-built tables)."""

    def __init__(self, nr, nphi, nz_local, dr, dphi, dz, R, comm, nz_per_rank=None):
        if not torch.cuda.is_available():
            raise RuntimeError("adi_thermal_fields_b200.slab: no CUDA device (there is no CPU fallback)")
        self.nr, self.nphi, self.nz = int(nr), int(nphi), int(nz_local)
        self.dr, self.dphi, self.dz, self.R = float(dr), float(dphi), float(dz), float(R)
        self.comm, self.rank, self.world = comm, comm.rank, comm.world
        self.nz_per_rank = [int(v) for v in (nz_per_rank if nz_per_rank is not None else [self.nz] * self.world)]
        if len(self.nz_per_rank) != self.world or self.nz_per_rank[self.rank] != self.nz:
            raise ValueError("nz_per_rank does not match the communicator / the local nz")
        self.L = _capi.load()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        h = C.c_void_p()
        _capi.check(self.L.adi_ctx_create(self.dev.index, C.byref(h)), "adi_ctx_create")
        self.ctx = h
        _capi.check(self.L.adi_cyl_bind(self.ctx, self.nr, self.nphi, self.nz, self.nz, self.dr, self.dphi, self.dz),
                    "adi_cyl_bind")
        _capi.check(self.L.adi_cyl_set_slab(self.ctx, self.rank, self.world, (C.c_int * self.world)(*self.nz_per_rank)),
                    "adi_cyl_set_slab")
        nl = self.nr * self.nphi
        self.y = torch.empty((2, nl), dtype=torch.float64, device=self.dev)
        self.y_all = torch.empty((self.world, 2, nl), dtype=torch.float64, device=self.dev)

    def launch_count(self):
        return int(self.L.adi_launch_count(self.ctx))


def adi_step_cyl(Tn, grid, mat, prm, robin_r, zbc, S=None, active=None, robin_inner=None, robin_void=None, out=None):
    """One backward-Euler step of the slab (collective over grid.comm): adi_step
    (adi3d_cyl_phi_v3.py:332-350) / adi_step_masked (quick_spiral_deposition_gif_v5.py:31-70) on device-resident
    local arrays (nr, nphi, nz_local).  Returns a new local array."""
    from . import adi3d_cyl_phi_v3 as gc
    if active is not None:
        ri, rv = robin_inner or robin_r, robin_void or robin_r
        p = gc._params(mat, prm, robin_r, zbc, T_void=float(rv.T_inf), T_inner=float(ri.T_inf))
    else:
        p = gc._params(mat, prm, robin_r, zbc)
    L, ctx = grid.L, grid.ctx
    st = torch.cuda.current_stream().cuda_stream
    if out is None:
        out = torch.empty_like(Tn)
    a = None if active is None else active.data_ptr()
    s = None if S is None else S.data_ptr()
    if grid.world == 1:
        _capi.check(L.adi_cyl_step(ctx, Tn.data_ptr(), out.data_ptr(), C.byref(p), a, s, st), "adi_cyl_step")
        return out
    _capi.check(L.adi_cyl_step_rphi(ctx, Tn.data_ptr(), out.data_ptr(), C.byref(p), a, s, st), "adi_cyl_step_rphi")
    _capi.check(L.adi_cyl_zsweep_reduce(ctx, out.data_ptr(), C.byref(p), grid.y.data_ptr(), st), "adi_cyl_zsweep_reduce")
    grid.comm.all_gather(grid.y_all, grid.y)
    _capi.check(L.adi_cyl_zsweep_finish(ctx, out.data_ptr(), C.byref(p), grid.y_all.data_ptr(), a, st),
                "adi_cyl_zsweep_finish")
    return out
