"""Minimal host-side containers for the inputs of the transport path.

The B200 algorithms accept either real OpenPNM objects (anything with
``['throat.conns']``, ``.Np``, ``.Nt``, ``.pores(label)``) or these light
stand-ins, which exist so the path can be driven on a machine where OpenPNM
itself is not installed (benchmarks, GPU tests).  Only what the solve reads is
reproduced: topology arrays and labels.  Geometry models, plotting, I/O, etc.
stay in OpenPNM.
"""
import numpy as np

__all__ = ["Network", "Cubic", "Phase"]


class _Store(dict):
    """dict of numpy arrays with OpenPNM's 'pore.' / 'throat.' key rules
    (openpnm/core/_base2.py:49-200, the subset the transport path touches):
    scalars broadcast to Np/Nt-long arrays, nested prefixes readable as dicts."""

    def _count(self, key):
        raise NotImplementedError

    def __setitem__(self, key, value):
        n = self._count(key)
        a = np.asarray(value)
        if a.ndim == 0 and n is not None:
            a = np.full(n, a[()])
        elif n is not None and a.shape[0] != n:
            raise ValueError(f"{key}: expected first dimension {n}, got {a.shape}")
        super().__setitem__(key, np.array(a))

    def __getitem__(self, key):
        try:
            return super().__getitem__(key)
        except KeyError:
            pre = key + "."
            sub = {k[len(pre):]: v for k, v in self.items() if k.startswith(pre)}
            if sub:
                return sub
            raise

    def __delitem__(self, key):
        if key in self.keys():
            return super().__delitem__(key)
        pre = key + "."
        hits = [k for k in self.keys() if k.startswith(pre)]
        if not hits:
            raise KeyError(key)
        for k in hits:
            super().__delitem__(k)


class Network(_Store):
    """Pore coordinates + throat connections (openpnm/network/_network.py)."""

    def __init__(self, coords=None, conns=None, Np=None):
        super().__init__()
        conns = np.asarray(conns, dtype=np.int64).reshape(-1, 2)
        # conns are forced upper-triangular on assignment (_network.py:123-128)
        conns = np.sort(conns, axis=1)
        if coords is not None:
            Np = np.asarray(coords).shape[0]
        if Np is None:
            Np = int(conns.max()) + 1 if conns.size else 0
        self._Np = int(Np)
        self._Nt = int(conns.shape[0])
        dict.__setitem__(self, "throat.conns", np.ascontiguousarray(conns))
        if coords is not None:
            dict.__setitem__(self, "pore.coords", np.asarray(coords, dtype=float))
        dict.__setitem__(self, "pore.all", np.ones(self._Np, dtype=bool))
        dict.__setitem__(self, "throat.all", np.ones(self._Nt, dtype=bool))

    def _count(self, key):
        if key.startswith("pore."):
            return self._Np
        if key.startswith("throat."):
            return self._Nt
        raise KeyError("keys must start with 'pore.' or 'throat.'")

    @property
    def Np(self):
        return self._Np

    @property
    def Nt(self):
        return self._Nt

    @property
    def Ps(self):
        return np.arange(self._Np)

    @property
    def Ts(self):
        return np.arange(self._Nt)

    def pores(self, labels="all"):
        if isinstance(labels, str):
            labels = [labels]
        mask = np.zeros(self._Np, dtype=bool)
        for lbl in labels:
            mask |= np.asarray(dict.__getitem__(self, "pore." + lbl.split("pore.")[-1]), dtype=bool)
        return np.flatnonzero(mask)

    def throats(self, labels="all"):
        if isinstance(labels, str):
            labels = [labels]
        mask = np.zeros(self._Nt, dtype=bool)
        for lbl in labels:
            mask |= np.asarray(dict.__getitem__(self, "throat." + lbl.split("throat.")[-1]),
                               dtype=bool)
        return np.flatnonzero(mask)

    def num_pores(self, labels="all"):
        return int(self.pores(labels).size)

    def num_throats(self, labels="all"):
        return int(self.throats(labels).size)

    @property
    def conns(self):
        return dict.__getitem__(self, "throat.conns")


class Cubic(Network):
    """Simple cubic lattice numbered exactly like `op.network.Cubic`.

    Pore index = x*Ny*Nz + y*Nz + z (openpnm/_skgraph/generators/_cubic.py:33-36);
    throats are the z-, then y-, then x-neighbour pairs (:40-42, 69-73); face
    labels left/right (x), front/back (y), bottom/top (z) and the xmin/xmax...
    aliases (_skgraph/generators/tools/_funcs.py:216-250)."""

    # neighbour offsets (dx, dy, dz) in the order the reference emits their throats
    # (_cubic.py:40-67): faces, then the four body diagonals, then the face diagonals
    _FACE_DIRS = ((0, 0, 1), (0, 1, 0), (1, 0, 0))
    _CORNER_DIRS = ((1, 1, 1), (1, 1, -1), (1, -1, 1), (-1, 1, 1))
    _EDGE_DIRS = ((0, 1, 1), (0, 1, -1), (1, 0, 1), (-1, 0, 1), (-1, -1, 0), (-1, 1, 0))

    def _dirs(self):
        try:
            return {6: self._FACE_DIRS, 14: self._FACE_DIRS + self._CORNER_DIRS,
                    18: self._FACE_DIRS + self._EDGE_DIRS,
                    20: self._EDGE_DIRS + self._CORNER_DIRS,
                    26: self._FACE_DIRS + self._CORNER_DIRS + self._EDGE_DIRS}[self.connectivity]
        except KeyError:
            raise Exception("Invalid connectivity. Must be 6, 14, 18, 20 or 26.") from None

    def __init__(self, shape, spacing=1.0, coords=True, lazy=False, connectivity=6):
        """connectivity: 6 (faces), 14 (+ body diagonals), 18 (+ face diagonals), 20
        (diagonals only) or 26, as `op.network.Cubic`.
        lazy=True keeps only the shape: 'throat.conns' (and 'throat.all') are generated
        when first read.  A slab-parallel run never reads them -- every rank derives the
        throats of its own planes from index arithmetic (openpnm_b200.dist) -- so a 400^3
        lattice costs no rank the 3 GB of the full list."""
        shape = np.array(shape, ndmin=1, dtype=int)
        if shape.size < 3:
            shape = np.concatenate((shape, np.ones(3 - shape.size, dtype=int)))
        nx, ny, nz = (int(s) for s in shape)
        n = nx * ny * nz
        self.shape = (nx, ny, nz)
        self.connectivity = int(connectivity)
        self.spacing = np.ones(3) * np.asarray(spacing, dtype=float)
        dirs = self._dirs()
        if lazy:
            _Store.__init__(self)
            self._Np = n
            self._Nt = sum(max(nx - abs(dx), 0) * max(ny - abs(dy), 0) * max(nz - abs(dz), 0)
                           for dx, dy, dz in dirs)
            dict.__setitem__(self, "pore.all", np.ones(n, dtype=bool))
        else:
            super().__init__(coords=None, conns=self._lattice_conns(), Np=n)
        if coords:
            ii, jj, kk = np.unravel_index(np.arange(n), (nx, ny, nz))
            c = (np.vstack((ii, jj, kk)).T + 0.5) * self.spacing
            dict.__setitem__(self, "pore.coords", c)

    def _lattice_conns(self):
        nx, ny, nz = self.shape
        if self.connectivity != 6:
            # per offset: the pores whose neighbour at that offset exists, in index order,
            # paired with that neighbour; each pair stored smaller index first
            # (_cubic.py:69-77)
            ax, ay, az = (np.arange(m, dtype=np.int64) for m in (nx, ny, nz))
            parts = []
            for dx, dy, dz in self._dirs():
                xs, ys, zs = (a[(a + d >= 0) & (a + d < m)]
                              for a, d, m in ((ax, dx, nx), (ay, dy, ny), (az, dz, nz)))
                tails = ((xs[:, None, None] * ny + ys[None, :, None]) * nz
                         + zs[None, None, :]).ravel()
                heads = tails + (dx * ny + dy) * nz + dz
                parts.append(np.stack((np.minimum(tails, heads), np.maximum(tails, heads)),
                                      axis=1))
            return np.concatenate(parts) if parts else np.zeros((0, 2), dtype=np.int64)
        idx = np.arange(nx * ny * nz, dtype=np.int64).reshape(nx, ny, nz)
        tails = np.concatenate((idx[:, :, :-1].ravel(), idx[:, :-1, :].ravel(),
                                idx[:-1, :, :].ravel()))
        heads = np.concatenate((idx[:, :, 1:].ravel(), idx[:, 1:, :].ravel(),
                                idx[1:, :, :].ravel()))
        conns = np.empty((tails.size, 2), dtype=np.int64)
        conns[:, 0] = tails
        conns[:, 1] = heads
        return conns

    def __missing__(self, key):
        # lazy lattice: materialise the throat arrays on first use
        if key == "throat.conns":
            dict.__setitem__(self, key, self._lattice_conns())
            return dict.__getitem__(self, key)
        if key == "throat.all":
            dict.__setitem__(self, key, np.ones(self._Nt, dtype=bool))
            return dict.__getitem__(self, key)
        raise KeyError(key)

    def __setitem__(self, key, value):
        if key == "throat.conns":
            self._conns_touched = True        # no longer the generator's lattice
        super().__setitem__(key, value)

    def _face(self, axis, last):
        """Sorted pore indices of one face, from index arithmetic (no Np-long temporaries)."""
        nx, ny, nz = self.shape
        ax, ay, az = (np.arange(m, dtype=np.int64) for m in (nx, ny, nz))
        if axis == 0:
            x = nx - 1 if last else 0
            return x * ny * nz + np.arange(ny * nz, dtype=np.int64)
        if axis == 1:
            y = ny - 1 if last else 0
            return (ax[:, None] * (ny * nz) + y * nz + az[None, :]).ravel()
        z = nz - 1 if last else 0
        return ((ax[:, None] * ny + ay[None, :]) * nz + z).ravel()

    _FACES = {"left": (0, False), "right": (0, True), "front": (1, False), "back": (1, True),
              "bottom": (2, False), "top": (2, True), "xmin": (0, False), "xmax": (0, True),
              "ymin": (1, False), "ymax": (1, True), "zmin": (2, False), "zmax": (2, True)}

    def pores(self, labels="all"):
        if isinstance(labels, str):
            labels = [labels]
        out = []
        rest = []
        for lbl in labels:
            key = lbl.split("pore.")[-1]
            if key in self._FACES and ("pore." + key) not in self.keys():
                out.append(self._face(*self._FACES[key]))
            else:
                rest.append(lbl)
        if rest:
            out.append(super().pores(rest))
        return np.unique(np.concatenate(out)) if out else np.array([], dtype=np.int64)


class Phase(_Store):
    """Bag of pore/throat properties bound to a network (openpnm/phase/_phase.py)."""

    def __init__(self, network, name="phase_01"):
        super().__init__()
        self.network = network
        self.name = name
        self.models = {}

    def add_model(self, propname, model, **kwargs):
        """Register a pore-scale model the way `phase.add_model` does
        (openpnm/core/_models.py): only the source-term models are evaluated
        by this package (on the device); `model` is the function or its name."""
        kwargs.pop("regen_mode", None)
        self.models[propname] = dict(model=model, **kwargs)

    def regenerate_models(self, propnames=None):
        """Run the callable models among `propnames` (all when None) in the given
        order; a dict result is stored as 'prop.key' entries, the way
        `ModelsMixin2.run_model` does (openpnm/core/_models.py:483-560).  Models
        named by a string are device kernels and are skipped here."""
        for prop in (list(self.models) if propnames is None else propnames):
            m = self.models.get(prop)
            if m is None or not callable(m["model"]):
                continue
            kwargs = {k: v for k, v in m.items() if k != "model"}
            vals = m["model"](self, **kwargs)
            if isinstance(vals, dict):
                for k, v in vals.items():
                    self[f"{prop}.{k}"] = v
            else:
                self[prop] = vals

    def _count(self, key):
        if key.startswith("pore."):
            return self.network.Np
        if key.startswith("throat."):
            return self.network.Nt
        raise KeyError("keys must start with 'pore.' or 'throat.'")
