"""TEST DOUBLE (tests only): a numpy particle store with the interface slab.py's exchange protocol uses, so
the protocol (cuts, migration, halo selection, bookkeeping) can run on CPU under gloo.  It does no SPH —
`step` just drifts owned particles — and is never imported by the product."""
import numpy as np
import torch


class NumpyStore:
    device = torch.device("cpu")

    def __init__(self, nsr: float, bounds):
        self.nsr = np.float32(nsr)
        self.bounds = bounds
        self.rec = np.zeros((0, 8), np.float32)
        self.ghost = np.zeros(0, bool)
        self.snapshots = []

    def configure(self, axis, own_lo, own_hi, layers, id_space, box_min, box_max):
        self.axis, self.own_lo, self.own_hi, self.layers = axis, own_lo, own_hi, layers

    def load(self, pos, vel, mass, ids):
        n = pos.shape[0]
        self.rec = np.zeros((n, 8), np.float32)
        self.rec[:, 0:3] = pos
        self.rec[:, 3] = 1.0 if mass is None else mass
        if vel is not None:
            self.rec[:, 4:7] = vel
        self.rec[:, 7] = np.asarray(ids, np.uint32).view(np.float32)
        self.ghost = np.zeros(n, bool)

    def _cells(self, rec):
        inv = np.float32(1.0) / self.nsr
        return np.floor(rec[:, self.axis] * inv).astype(np.int64)

    def exchange_pack(self, cuts, me, buf):
        """Same contract as sphb_slab_exchange_pack: groups [owned by r][ghosts for r] for r = 0..G-1."""
        G = len(cuts) - 1
        rec = self.rec[~self.ghost]
        cells = self._cells(rec)
        owner = np.clip(np.searchsorted(cuts, cells, side="right") - 1, 0, G - 2 + 1)
        L = self.layers
        lo_edge = np.asarray(cuts)[owner].astype(np.int64)
        hi_edge = np.asarray(cuts)[owner + 1].astype(np.int64)
        ghost_lo = np.where((owner > 0) & (cells < lo_edge + L), owner - 1, -1)
        ghost_hi = np.where((owner < G - 1) & (cells >= hi_edge - L), owner + 1, -1)
        counts = np.zeros(2 * G, np.uint64)
        out = buf.numpy()
        k = 0
        for r in range(G):
            own = rec[owner == r]
            counts[2 * r] = len(own)
            if r != me:
                out[k:k + len(own)] = own
                k += len(own)
            gh = rec[(ghost_lo == r) | (ghost_hi == r)].copy()
            gh[:, 7] = (gh[:, 7].view(np.uint32) | np.uint32(0x80000000)).view(np.float32)
            counts[2 * r + 1] = len(gh)
            out[k:k + len(gh)] = gh
            k += len(gh)
        self.rec = rec[owner == me].copy()
        self.ghost = np.zeros(len(self.rec), bool)
        return counts

    # two-phase form (sphb_slab_exchange_count / _split): the routing is computed once, by the count phase
    def exchange_count(self, cuts, me, d_counts):
        self._staged = torch.zeros((max(3 * len(self.rec), 16), 8), dtype=torch.float32)
        counts = self.exchange_pack(cuts, me, self._staged)
        d_counts.copy_(torch.from_numpy(counts.astype(np.int32)))

    def exchange_split(self, cuts, me, counts, buf):
        n_out = int(np.asarray(counts, np.int64).sum() - int(counts[2 * me]))
        buf[:n_out].copy_(self._staged[:n_out])

    def append(self, buf, count, ghost=None):
        if not count:
            return
        new = buf.numpy()[:count].copy()
        word = new[:, 7].view(np.uint32)
        flag = (word & np.uint32(0x80000000)) != 0 if ghost is None else np.full(count, bool(ghost))
        new[:, 7] = (word & np.uint32(0x7FFFFFFF)).view(np.float32)
        self.rec = np.concatenate([self.rec, new])
        self.ghost = np.concatenate([self.ghost, flag])

    def ids(self, ghost):
        return np.sort(self.rec[self.ghost == ghost, 7].view(np.uint32))

    def step(self, dt):
        self.snapshots.append((self.ids(False), self.ids(True)))
        own = ~self.ghost
        self.rec[own, 0:3] += self.rec[own, 4:7] * np.float32(dt)
        lo = np.array(self.bounds[0::2], np.float32); hi = np.array(self.bounds[1::2], np.float32)
        p = self.rec[own, 0:3]
        v = self.rec[own, 4:7]
        below, above = p < lo, p > hi
        v[below | above] *= np.float32(-0.8)
        self.rec[own, 0:3] = np.minimum(np.maximum(p, lo), hi)
        self.rec[own, 4:7] = v

    def synchronize(self):
        pass

    def download(self, **kw):
        r = self.rec[~self.ghost]
        return {"ids": r[:, 7].view(np.uint32).copy(), "pos": r[:, 0:3].copy(), "vel": r[:, 4:7].copy()}

    @property
    def size(self):
        return len(self.rec)


def expected_sets(pos, cuts, axis, nsr, layers):
    """Brute force: for every rank the ids it must own and the ids it must hold as ghosts."""
    inv = np.float32(1.0) / np.float32(nsr)
    cells = np.floor(pos[:, axis].astype(np.float32) * inv).astype(np.int64)
    owner = np.clip(np.searchsorted(cuts, cells, side="right") - 1, 0, len(cuts) - 2)
    out = []
    for d in range(len(cuts) - 1):
        own = np.flatnonzero(owner == d)
        lo, hi = int(cuts[d]), int(cuts[d + 1])
        gh = np.flatnonzero(((owner == d - 1) & (cells >= lo - layers)) | ((owner == d + 1) & (cells < hi + layers)))
        out.append((own.astype(np.uint32), gh.astype(np.uint32)))
    return out
