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

    def extract_migrants(self, cuts, me, buf):
        rec = self.rec[~self.ghost]
        dest = np.clip(np.searchsorted(cuts, self._cells(rec), side="right") - 1, 0, len(cuts) - 2)
        counts = np.bincount(dest, minlength=len(cuts) - 1).astype(np.uint64)
        out = buf.numpy()
        k = 0
        for d in range(len(cuts) - 1):
            if d == me:
                continue
            sel = rec[dest == d]
            out[k:k + len(sel)] = sel
            k += len(sel)
        self.rec = rec[dest == me].copy()
        self.ghost = np.zeros(len(self.rec), bool)
        return counts

    def extract_halo(self, side, buf):
        rec = self.rec[~self.ghost]
        c = self._cells(rec)
        lo, hi = (self.own_lo, self.own_lo + self.layers) if side == 0 else (self.own_hi - self.layers, self.own_hi)
        sel = rec[(c >= lo) & (c < hi)]
        buf.numpy()[: len(sel)] = sel
        return len(sel)

    def append(self, buf, count, ghost):
        if count:
            self.rec = np.concatenate([self.rec, buf.numpy()[:count].copy()])
            self.ghost = np.concatenate([self.ghost, np.full(count, bool(ghost))])

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
