"""Slab domain decomposition across GPUs: host-side planning and the per-step exchange protocol.

NEW functionality — the reference is a single-process CPU program.  One process per
GPU (torch.distributed, NCCL over NVLink); rank d owns the reference cells [cuts[d], cuts[d+1]) along one
axis.  Per step:

  1. exchange    ONE round: every owned particle is routed to the rank that owns its cell now (migration —
                 general, not just ±1 neighbours) and, flagged as ghost, to each adjacent rank whose two
                 halo layers contain the cell.  Group sizes are all-gathered (16 B x ranks^2), the 32-byte
                 records travel in one all_to_all_single.  With two halo layers the density of the first
                 ghost layer is recomputed locally, so no mid-step exchange of densities is needed.
  2. local step  sphb_step on owned + ghosts; ghosts are neither advanced nor kept

All packing / unpacking runs in CUDA kernels of libsphb (csrc/slab.cu); this module only plans the cuts
and moves device buffers.  Because every local step re-sorts owned + ghost particles by (cell, global id),
an N-GPU run reproduces the 1-GPU run bit for bit in strict math mode (tests/test_slab_*.py).

The exchange logic is written against a small "store" interface so that it can be exercised on CPU with
gloo (tests/ supplies a numpy store as a test double); the product store is GpuStore below.
"""
from __future__ import annotations

import numpy as np
import torch

REC = 8                    # floats per exchange record {x, y, z, m, vx, vy, vz, id}
OPEN_LO, OPEN_HI = -(1 << 28), (1 << 28)   # open-ended first / last slab


def axis_cells(pos: np.ndarray, axis: int, nsr: float) -> np.ndarray:
    """Reference cell index along `axis`: (int)floorf(p * (1.0f / nsr)) in fp32 (spatial_hash.h:30-36)."""
    inv = np.float32(1.0) / np.float32(nsr)
    return np.floor(pos[:, axis].astype(np.float32) * inv).astype(np.int64)


def plan_cuts(cells: np.ndarray, nranks: int, min_width: int = 2) -> np.ndarray:
    """Cut the occupied cell range into `nranks` slabs of whole cells with near-equal particle counts.

    Returns int32[nranks + 1]; the end slabs are open-ended so every cell maps to exactly one rank.
    Every slab is at least `min_width` cells wide (the halo depth), which keeps halos nearest-neighbour."""
    lo, hi = int(cells.min()), int(cells.max()) + 1
    if hi - lo < nranks * min_width:
        raise ValueError(f"{hi - lo} cells along the slab axis cannot host {nranks} slabs of >= {min_width} cells")
    hist = np.bincount(cells - lo, minlength=hi - lo).astype(np.int64)
    cum = np.concatenate([[0], np.cumsum(hist)])
    total = cum[-1]
    inner = []
    prev = lo
    for d in range(1, nranks):
        target = total * d / nranks
        c = lo + int(np.searchsorted(cum, target, side="left"))
        # choose the closer of the two neighbouring cell faces
        if c > lo and abs(cum[c - lo - 1] - target) < abs(cum[min(c - lo, hi - lo)] - target):
            c -= 1
        c = max(c, prev + min_width)
        c = min(c, hi - (nranks - d) * min_width)
        inner.append(c)
        prev = c
    return np.array([OPEN_LO] + inner + [OPEN_HI], dtype=np.int32)


def rank_of_cells(cuts: np.ndarray, cells: np.ndarray) -> np.ndarray:
    return np.clip(np.searchsorted(cuts, cells, side="right") - 1, 0, len(cuts) - 2)


class GpuStore:
    """One sphb context on one GPU, seen through the store interface the exchange protocol uses."""

    def __init__(self, pkg, capacity: int, device_index: int, params: dict, strict: bool = False, stream=None, options=None):
        self.pkg = pkg
        self.device = torch.device("cuda", device_index)
        self.ctx = pkg.Context(capacity, device_index)
        capi = pkg.capi
        self.ctx.set_option(capi.OPT_MATH_MODE, capi.MATH_STRICT if strict else capi.MATH_FAST)
        for k, v in (options or {}).items():
            self.ctx.set_option(k, v)
        if stream is not None:
            self.ctx.set_stream(stream)
        self.ctx.set_params(params)
        self.params = dict(params)
        self.capacity = capacity

    def configure(self, axis, own_lo, own_hi, layers, id_space, box_min, box_max):
        self.ctx.set_slab(axis, own_lo, own_hi, layers, id_space, box_min, box_max)

    def load(self, pos, vel, mass, ids):
        self.ctx.upload_ids(pos, vel, mass, ids)

    def exchange_pack(self, cuts, me, buf: torch.Tensor):
        """counts[2r] = records owned by rank r (kept in place for r = me), counts[2r+1] = ghosts for rank r."""
        return self.ctx.slab_exchange_pack(cuts, me, buf.data_ptr(), buf.shape[0])

    def exchange_count(self, cuts, me, d_counts: torch.Tensor):
        """Two-phase form, phase 1 (no host synchronisation): group sizes -> int32 device tensor."""
        self.ctx.slab_exchange_count(cuts, me, d_counts.data_ptr())

    def exchange_split(self, cuts, me, counts, buf: torch.Tensor):
        """Phase 2: move the records given this rank's group sizes (host)."""
        self.ctx.slab_exchange_split(cuts, me, counts, buf.data_ptr(), buf.shape[0])

    def append(self, buf: torch.Tensor, count: int, ghost=None):
        """ghost=None: each record carries its own ghost flag (id bit 31)."""
        if count:
            self.ctx.slab_append(buf.data_ptr(), count, ghost)

    def step(self, dt):
        self.ctx.step(dt)

    def cfl_state(self):
        return self.ctx.get_cfl_state()

    def set_cfl_state(self, max_v2, a0):
        self.ctx.set_cfl_state(max_v2, a0)

    def synchronize(self):
        self.ctx.synchronize()

    def download(self, **kw):
        return self.ctx.slab_download(**kw)

    @property
    def size(self):
        return self.ctx.size

    def close(self):
        self.ctx.close()


class SlabRank:
    """Per-rank state of the decomposition: its slab, its store and its exchange buffers."""

    def __init__(self, store, rank: int, cuts: np.ndarray, axis: int, layers: int, id_space: int, box_min, box_max,
                 exchange_capacity: int):
        self.store, self.rank, self.cuts, self.axis, self.layers = store, rank, np.asarray(cuts, np.int32), axis, layers
        self.nranks = len(cuts) - 1
        for d in range(self.nranks):
            if self.cuts[d + 1] - self.cuts[d] < layers:
                raise ValueError("slab thinner than the halo depth")
        store.configure(axis, int(cuts[rank]), int(cuts[rank + 1]), layers, id_space, box_min, box_max)
        dev = store.device
        self.send = torch.empty((exchange_capacity, REC), dtype=torch.float32, device=dev)
        self.recv = torch.empty((exchange_capacity, REC), dtype=torch.float32, device=dev)
        # two-phase exchange: my group sizes, the gathered table and its pinned host mirror
        self.d_counts = torch.zeros(2 * self.nranks, dtype=torch.int32, device=dev)
        self.d_table = torch.zeros((self.nranks, 2 * self.nranks), dtype=torch.int32, device=dev)
        self.h_table = torch.zeros((self.nranks, 2 * self.nranks), dtype=torch.int32)
        if dev.type == "cuda":
            self.h_table = self.h_table.pin_memory()
        self.stats = {"migrants_sent": 0, "halo_sent": 0, "steps": 0}
        self.params = getattr(store, "params", None)      # needed only for adaptive timesteps

    def load_initial(self, pos, vel, mass, nsr: float):
        """Every rank sees the same synthetic scene and keeps the particles of its own slab (global id = index)."""
        cells = axis_cells(pos, self.axis, nsr)
        mine = np.flatnonzero(rank_of_cells(self.cuts, cells) == self.rank)
        self.store.load(pos[mine], None if vel is None else vel[mine], None if mass is None else mass[mine], mine.astype(np.uint32))
        return mine

    def pack(self):
        """Route all owned particles; returns (per-destination send sizes, raw 2G counts with 'kept' zeroed)."""
        counts = np.array(self.store.exchange_pack(self.cuts, self.rank, self.send), dtype=np.int64)
        counts[2 * self.rank] = 0                                   # kept in place, not sent
        self.stats["migrants_sent"] += int(counts[0::2].sum())
        self.stats["halo_sent"] += int(counts[1::2].sum())
        return counts


def cfl_timestep(params: dict, max_v2, a0) -> np.float32:
    """SPHEngine::compute_cfl_timestep (reference sph_engine.cpp:312-333) from globally reduced inputs, in the
    same fp32 operation order as the device kernel k_cfl_dt (so every rank derives the identical dt)."""
    f = np.float32
    cfl, h, ts = f(params["CFL_factor"]), f(params["smoothing_length"]), f(params["timestep"])
    max_velocity = np.sqrt(f(max_v2))
    dt_cfl = f(cfl * h) / f(max_velocity + f(1e-6))
    a = f(a0[0]) * f(a0[0]) + f(a0[1]) * f(a0[1])
    a = np.sqrt(f(f(a) + f(a0[2]) * f(a0[2])))
    dt_force = f(cfl * np.sqrt(f(h / f(a + f(1e-6)))))
    m = dt_cfl
    if dt_force < m:
        m = dt_force
    if ts < m:
        m = ts
    return f(m)


def reduce_cfl_state(states):
    """states: [(max_v2, a0, fresh)] of every rank → (global max_v2, a0 of the rank that last advanced id 0 or None)."""
    v2 = max(np.float32(s[0]) for s in states)
    fresh = [s[1] for s in states if s[2]]
    return v2, (fresh[0] if fresh else None)


def _splits(table: np.ndarray, me: int):
    """table[src, 2*dst + {0 owned, 1 ghost}] → (records I send to each rank, records I receive from each rank)."""
    G = table.shape[0]
    send = [int(table[me, 2 * d] + table[me, 2 * d + 1]) for d in range(G)]
    recv = [int(table[src, 2 * me] + table[src, 2 * me + 1]) for src in range(G)]
    return send, recv


def step_local(ranks: list[SlabRank], dt: float):
    """All ranks inside ONE process (several contexts, possibly on one GPU): the same protocol with device
    copies instead of NCCL.  Used by the single-GPU tests of the multi-GPU path."""
    G = len(ranks)
    table = np.stack([r.pack() for r in ranks])
    for r in ranks:
        r.store.synchronize()
    offs = []
    for src in range(G):
        send, _ = _splits(table, src)
        offs.append(np.concatenate([[0], np.cumsum(send)]))
    for dst in range(G):
        for src in range(G):
            n = int(offs[src][dst + 1] - offs[src][dst])
            if n:
                tmp = ranks[src].send[offs[src][dst]: offs[src][dst] + n].to(ranks[dst].store.device).contiguous()
                if tmp.is_cuda:
                    torch.cuda.synchronize()
                ranks[dst].store.append(tmp, n, None)
                ranks[dst].store.synchronize()
    if dt <= 0.0:                                            # adaptive: global CFL inputs, identical dt everywhere
        states = [r.store.cfl_state() for r in ranks]
        v2, a0 = reduce_cfl_state(states)
        for r, st in zip(ranks, states):
            r.a0 = a0 if a0 is not None else getattr(r, "a0", np.zeros(3, np.float32))
        dt = float(cfl_timestep(ranks[0].params, v2, ranks[0].a0))
    for r in ranks:
        r.store.step(dt)
        r.stats["steps"] += 1
    return dt


def step_distributed(r: SlabRank, dt: float, group=None):
    """One process per GPU: the protocol over torch.distributed (NCCL on GPUs, gloo in the CPU tests).
    One host synchronisation per step (the gathered table of group sizes) and two collectives (all_gather of 2G
    integers, all_to_all of the records)."""
    import torch.distributed as dist

    G, me, dev = r.nranks, r.rank, r.store.device
    if hasattr(r.store, "exchange_count"):
        # ONE host synchronisation: the routing count, the all-gather of the group sizes and the copy of the
        # gathered table to pinned memory are enqueued back to back; the host then waits once, and everything
        # after it (record split, all_to_all, append, the local step) is enqueued without further waiting
        r.store.exchange_count(r.cuts, me, r.d_counts)
        dist.all_gather_into_tensor(r.d_table.view(-1), r.d_counts, group=group)
        if dev.type == "cuda":    # kernel-written read-back: a copy would queue behind a bulk read-back in flight on the copy engine
            r.store.ctx.read_small(r.d_table.data_ptr(), r.h_table.data_ptr(), r.d_table.numel() * r.d_table.element_size())
        else:
            r.h_table.copy_(r.d_table, non_blocking=True)
        if dev.type == "cuda":
            torch.cuda.current_stream(dev).synchronize()
        table = r.h_table.numpy().astype(np.int64)
        r.store.exchange_split(r.cuts, me, table[me], r.send)
        table[me, 2 * me] = 0                                        # kept in place, not sent
        r.stats["migrants_sent"] += int(table[me, 0::2].sum())
        r.stats["halo_sent"] += int(table[me, 1::2].sum())
    else:                                                            # stores without the two-phase form (CPU test double)
        counts = r.pack()
        mine = torch.from_numpy(counts).to(dev)
        table = torch.empty((G, 2 * G), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(table.view(-1), mine, group=group)
        table = table.cpu().numpy()
    send, recv = _splits(table, me)
    n_out, n_in = sum(send), sum(recv)
    if n_in > r.recv.shape[0]:
        raise RuntimeError(f"rank {me}: {n_in} incoming records exceed the exchange buffer ({r.recv.shape[0]})")
    dist.all_to_all_single(r.recv[:n_in].view(-1), r.send[:n_out].view(-1), output_split_sizes=[c * REC for c in recv],
                           input_split_sizes=[c * REC for c in send], group=group)
    r.store.append(r.recv, n_in, None)
    if dt <= 0.0:                                            # adaptive: all-reduce the CFL inputs
        v2, a0, fresh = r.store.cfl_state()
        buf = torch.tensor([float(v2), float(a0[0]) if fresh else 0.0, float(a0[1]) if fresh else 0.0,
                            float(a0[2]) if fresh else 0.0, 1.0 if fresh else 0.0], dtype=torch.float32, device=dev)
        vmax = buf[:1].clone()
        dist.all_reduce(vmax, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(buf[1:], op=dist.ReduceOp.SUM, group=group)     # exactly one rank is fresh: the sum IS its a0
        host = buf.cpu().numpy()
        if host[4] > 0:
            r.a0 = host[1:4].astype(np.float32)
        elif not hasattr(r, "a0"):
            r.a0 = np.zeros(3, np.float32)
        dt = float(cfl_timestep(r.params, np.float32(vmax.item()), r.a0))
    r.store.step(dt)
    r.stats["steps"] += 1
    return dt


def gather_by_id(parts: list[dict], n_total: int) -> dict:
    """Merge per-rank owned-particle downloads into insertion-order arrays."""
    out = {}
    for key in ("pos", "vel", "rho", "P", "acc"):
        if key in parts[0]:
            shape = (n_total, 3) if parts[0][key].ndim == 2 else (n_total,)
            out[key] = np.zeros(shape, np.float32)
    seen = np.zeros(n_total, np.int32)
    for p in parts:
        ids = p["ids"].astype(np.int64)
        seen[ids] += 1
        for key in out:
            out[key][ids] = p[key]
    out["owners"] = seen
    return out
