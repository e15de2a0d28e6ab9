"""Lattice-T split with halo exchange (SURVEY §8e "secondary partitioning", BASELINE.json config 5): rank r owns the
global time-slices [r*Tl, (r+1)*Tl).  Spatial displacements need no communication; a displacement of length k in
+-t needs k time-slices of every eigenvector from the neighbouring rank.

Replaces, for a partitioned t-direction, QUDA's per-hop `exchangeGhost` (/root/reference/lib/contract_wrappers.cu:166-174,
one nFace=1 spinor halo per hop and eigenvector) and the extended gauge field with `exchangeExtendedGhost`
(lib/displace.cpp:104-134): here each eigenvector batch is extended ONCE by H slices on both sides (H = longest
t-displacement, rounded up to even so that the even/odd parity of a site is the same in local, extended and global
coordinates), the fused kernel then runs on the extended lattice (Lt_ext = Tl + 2H) as if it were periodic, and only
the interior is kept: for interior sites no shift of length <= H wraps, and the plus-direction loops of the lower halo
are valid too, which is what the minus-from-plus derivation of the interior needs.  In the even/odd site-major layout
a time-slice is one contiguous chunk per parity, so every halo is two contiguous blocks: they travel as NCCL P2P
(NVLink) send/recv pairs, batched per eigenvector batch.  The gauge field is replicated (<= 6.1 GB at 48^3x96), each
rank slices its extended slab once.  The result needs no reduction: each rank owns its time-slices, the projected
buffers are all-gathered (the reference's COMM_TIME gather, lib/loop_mugiq.cpp:420-424).
"""
import numpy as np
import torch


def _as_batch(views):
    """[nb, ...] strided view over equally spaced, equally shaped contiguous tensors of one storage, else None."""
    v0 = views[0]
    if len(views) == 1:
        return v0.unsqueeze(0)
    step = (views[1].data_ptr() - v0.data_ptr()) // v0.element_size()
    if step <= 0 or not all(v.is_contiguous() and v.untyped_storage().data_ptr() == v0.untyped_storage().data_ptr() and
                            v.data_ptr() == v0.data_ptr() + k * step * v0.element_size() for k, v in enumerate(views)):
        return None
    return torch.as_strided(v0, (len(views),) + tuple(v0.shape), (step,) + tuple(v0.stride()), v0.storage_offset())


class TSplit:
    def __init__(self, L_global, rank, world, max_t_disp):
        Lx, Ly, Lz, T = (int(x) for x in L_global)
        if T % world:
            raise ValueError(f"T = {T} is not divisible by {world} ranks")
        self.rank, self.world = int(rank), int(world)
        self.Tl = T // world
        if self.Tl % 2:
            raise ValueError(f"local T = {self.Tl} must be even (even/odd site order)")
        H = int(max_t_disp)
        H += H & 1
        if H > self.Tl:
            raise ValueError(f"t-displacements of length {max_t_disp} exceed the local time extent {self.Tl}")
        self.H = H
        self.L_global = (Lx, Ly, Lz, T)
        self.L_loc = (Lx, Ly, Lz, self.Tl)
        self.L_ext = (Lx, Ly, Lz, self.Tl + 2 * H)
        self.V3h = Lx * Ly * Lz // 2
        self.t0 = self.rank * self.Tl  # first global time-slice owned
        self.peer = None               # set by attach_peers: halos travel as direct NVLink writes into the neighbours' slabs

    # ---- NVLink peer mode ------------------------------------------------------------------------------------------------
    def attach_peers(self, slabs, up_ptr, dn_ptr, mode=0, group=None):
        """`slabs`: this rank's eigenvectors [nvec, V4_ext, 12] (complex128) in ONE allocation made by ops.PeerBuffer;
        `up_ptr` / `dn_ptr`: the same allocation of rank+1 / rank-1 mapped into this process (ops.peer_open; the rank's
        own pointer when world == 1).  From now on begin_extend on views of `slabs` writes the boundary slices of the batch
        straight into the neighbours' halo slices (mode 0: copy engines, mode 1: SM push kernel) on a high-priority side
        stream, followed by a one-element all-reduce that tells every rank its halos have landed."""
        if slabs.dim() != 3 or slabs.shape[1] != 2 * (self.Tl + 2 * self.H) * self.V3h or not slabs.is_contiguous():
            raise ValueError("attach_peers: slabs must be a contiguous [nvec, V4_ext, 12] tensor")
        self.peer = {"slabs": slabs, "up": int(up_ptr), "dn": int(dn_ptr), "mode": int(mode), "group": group,
                     "stream": torch.cuda.Stream(device=slabs.device, priority=-1),
                     "flag": torch.zeros(1, dtype=torch.float32, device=slabs.device)}

    def _begin_extend_peer(self, vectors, lower, upper):
        from . import ops
        import torch.distributed as dist
        pr = self.peer
        slabs = pr["slabs"]
        vec_bytes = slabs.shape[1] * slabs.shape[2] * slabs.element_size()
        first = (vectors[0].data_ptr() - slabs.data_ptr()) // vec_bytes
        nb = len(vectors)
        # consecutive vectors of the attached slabs (checked once per distinct batch; the views are cached with it)
        key = (vectors[0].data_ptr(), vectors[-1].data_ptr(), nb)
        cached = pr.setdefault("batches", {}).get(key)
        if cached is None:
            if (vectors[0].data_ptr() - slabs.data_ptr()) % vec_bytes or any(
                    v.data_ptr() != slabs.data_ptr() + (first + k) * vec_bytes for k, v in enumerate(vectors)):
                raise ValueError("begin_extend (peer mode): the batch must be consecutive vectors of the attached slabs")
            cached = [v.reshape(slabs.shape[1], -1) for v in vectors]
            pr["batches"][key] = cached
        H, Tl, Lt_ext = self.H, self.Tl, self.Tl + 2 * self.H
        site_bytes = 12 * slabs.element_size()
        cur = torch.cuda.current_stream(slabs.device)
        pr["stream"].wait_stream(cur)  # the interiors are final
        with torch.cuda.stream(pr["stream"]):
            if upper:  # my first `upper` interior slices are the upper halo of the rank below
                ops.halo_push_t(pr["dn"], slabs.data_ptr(), first, nb, Lt_ext, self.V3h, H, H + Tl, upper, pr["mode"], site_bytes)
            if lower:  # my last `lower` interior slices are the lower halo of the rank above
                ops.halo_push_t(pr["up"], slabs.data_ptr(), first, nb, Lt_ext, self.V3h, H + Tl - lower, H - lower, lower,
                                pr["mode"], site_bytes)
            if self.world > 1 and (lower or upper):
                dist.all_reduce(pr["flag"], group=pr["group"])  # every rank's pushes precede its contribution
            ev = torch.cuda.Event()
            ev.record(pr["stream"])
        return {"peer_event": ev, "views": cached}

    # ---- views: a full-lattice even/odd array [..., V4, C...] as [..., parity, t, V3/2, C...] ---------------------------
    def _view(self, a, Lt, site_dim):
        shp = list(a.shape)
        return a.reshape(shp[:site_dim] + [2, Lt, self.V3h] + shp[site_dim + 1:])

    def global_slab(self, field_global, site_dim=0):
        """Extended slab of this rank cut out of a GLOBAL even/odd field (numpy or torch), periodic in t.  Used for the
        replicated gauge field and by single-process tests."""
        v = self._view(field_global, self.L_global[3], site_dim)
        T = self.L_global[3]
        ts = [(self.t0 - self.H + i) % T for i in range(self.Tl + 2 * self.H)]
        idx = torch.as_tensor(ts, device=v.device) if isinstance(v, torch.Tensor) else np.asarray(ts)
        ext = v.index_select(site_dim + 1, idx) if isinstance(v, torch.Tensor) else np.take(v, idx, axis=site_dim + 1)
        shp = list(field_global.shape)
        shp[site_dim] = 2 * (self.Tl + 2 * self.H) * self.V3h
        return ext.reshape(shp)

    def interior(self, field_ext, site_dim=0):
        """Interior (owned time-slices) of an extended even/odd array, as a contiguous local-lattice array."""
        v = self._view(field_ext, self.Tl + 2 * self.H, site_dim)
        sl = [slice(None)] * v.dim()
        sl[site_dim + 1] = slice(self.H, self.H + self.Tl)
        out = v[tuple(sl)]
        shp = list(field_ext.shape)
        shp[site_dim] = 2 * self.Tl * self.V3h
        return out.reshape(shp)

    def begin_extend(self, vectors, group=None, device=None, lower=True, upper=True):
        """Starts the extension of a batch of eigenvectors (sequence of [V4_loc, 12] tensors in local even/odd order, or
        one [nb, V4_loc, 12] tensor): the interiors are written into the extended buffer and the halo send/recv pairs are
        posted (asynchronously: they overlap whatever is launched before finish_extend).  `lower` / `upper`: how many halo
        slices below / above the interior the kernels will read (True = all H; LoopPlan.t_halo(): plus-t loops of length
        k read k slices above the interior, directly computed minus-t loops below it; H itself is rounded up to even for
        the parity bookkeeping, but only the slices that are read are sent and filled).  Every rank must pass the same
        values.  Returns a handle."""
        import torch.distributed as dist
        nb = len(vectors)
        H, Tl = self.H, self.Tl
        v0 = vectors[0]
        device = device if device is not None else v0.device
        ncomp = 12
        lower = H if lower is True else min(int(lower), H)
        upper = H if upper is True else min(int(upper), H)
        if self.peer is not None and v0.numel() == 2 * (Tl + 2 * H) * self.V3h * ncomp:
            return self._begin_extend_peer(vectors, lower, upper)
        top = bot = None
        if v0.numel() == 2 * (Tl + 2 * H) * self.V3h * ncomp and v0.device == torch.device(device) and H > 0:
            # the caller already stores its slab in the extended layout (halo slices allocated, interior filled): only the
            # halos move, nothing is copied
            views = [v.reshape(2, Tl + 2 * H, self.V3h, ncomp) for v in vectors]
            batch = _as_batch(views)  # one strided view when the fields are slices of one allocation: 2 copies, not 2*nb
            h = {"ext": None, "views": views, "batch": batch, "reqs": [], "from_dn": None, "from_up": None,
                 "lower": lower, "upper": upper}
            if batch is not None:
                if lower:
                    top = batch[:, :, H + Tl - lower:H + Tl].contiguous()
                if upper:
                    bot = batch[:, :, H:H + upper].contiguous()
            else:
                if lower:
                    top = torch.stack([v[:, H + Tl - lower:H + Tl] for v in views])
                if upper:
                    bot = torch.stack([v[:, H:H + upper] for v in views])
        else:
            ext = torch.empty((nb, 2, Tl + 2 * H, self.V3h, ncomp), dtype=v0.dtype, device=device)
            for k in range(nb):
                ext[k, :, H:H + Tl].copy_(vectors[k].reshape(2, Tl, self.V3h, ncomp), non_blocking=True)
            h = {"ext": ext, "views": None, "reqs": [], "from_dn": None, "from_up": None, "lower": lower, "upper": upper}
            if lower:
                top = ext[:, :, H + Tl - lower:H + Tl].contiguous()  # last owned slices: the LOWER halo of the rank above
            if upper:
                bot = ext[:, :, H:H + upper].contiguous()            # first owned slices: the UPPER halo of the rank below
        if lower or upper:
            if self.world == 1:
                h["from_dn"], h["from_up"] = top, bot
            else:
                up, dn = (self.rank + 1) % self.world, (self.rank - 1) % self.world
                ops = []
                if lower:
                    h["from_dn"] = torch.empty_like(top)
                    ops += [dist.P2POp(dist.isend, top, up, group), dist.P2POp(dist.irecv, h["from_dn"], dn, group)]
                if upper:
                    h["from_up"] = torch.empty_like(bot)
                    ops += [dist.P2POp(dist.isend, bot, dn, group), dist.P2POp(dist.irecv, h["from_up"], up, group)]
                h["reqs"] = dist.batch_isend_irecv(ops)
                h["keep"] = (top, bot)  # the send buffers must outlive the transfers
        return h

    def finish_extend(self, h):
        """Waits for the halos of a begin_extend handle and returns the extended batch ([nb, V4_ext, 12], or the list of
        the caller's own extended vectors when they were extended in place)."""
        H, Tl = self.H, self.Tl
        if "peer_event" in h:  # the neighbours wrote the halos in place
            torch.cuda.current_stream().wait_event(h["peer_event"])
            return h["views"]
        for req in h["reqs"]:
            req.wait()
        lo, up = h["lower"], h["upper"]
        if h["views"] is not None:
            if h["batch"] is not None:
                if h["from_dn"] is not None:
                    h["batch"][:, :, H - lo:H] = h["from_dn"]
                if h["from_up"] is not None:
                    h["batch"][:, :, H + Tl:H + Tl + up] = h["from_up"]
            else:
                for k, v in enumerate(h["views"]):
                    if h["from_dn"] is not None:
                        v[:, H - lo:H] = h["from_dn"][k]
                    if h["from_up"] is not None:
                        v[:, H + Tl:H + Tl + up] = h["from_up"][k]
            return [v.reshape(2 * (Tl + 2 * H) * self.V3h, -1) for v in h["views"]]
        ext = h["ext"]
        if h["from_dn"] is not None:
            ext[:, :, H - lo:H] = h["from_dn"]
        if h["from_up"] is not None:
            ext[:, :, H + Tl:H + Tl + up] = h["from_up"]
        return ext.reshape(ext.shape[0], 2 * (Tl + 2 * H) * self.V3h, -1)

    def exchange_loop_halo(self, dataPosExt, slots, group=None, depth=None):
        """Fills the LOWER halo slices of the given loop slots of the extended position-space buffer
        [nLoop, 16, V4_ext] with the top interior slices of the rank below (periodic).  A minus-t loop derived from its
        plus-t partner reads the partner at x - k t: for the first k interior slices that is the neighbour's interior,
        which the neighbour has computed anyway - one exchange of 16*H*V3 complex per loop and run instead of an
        eigenvector halo (and its contraction) per eigenvector."""
        import torch.distributed as dist
        H, Tl = self.H, self.Tl
        d = H if depth is None else min(int(depth), H)  # slices the derived loops read below the interior
        if d == 0 or len(slots) == 0:
            return
        v = dataPosExt.reshape(dataPosExt.shape[0], dataPosExt.shape[1], 2, Tl + 2 * H, self.V3h)
        slots = sorted(int(x) for x in slots)
        # runs of consecutive slots are plain slices (no index tensor, no host-device synchronisation)
        runs, a = [], 0
        for k in range(1, len(slots) + 1):
            if k == len(slots) or slots[k] != slots[k - 1] + 1:
                runs.append((slots[a], slots[k - 1] + 1))
                a = k
        send = torch.cat([v[a:b, :, :, H + Tl - d:H + Tl] for a, b in runs]).contiguous()
        if self.world == 1:
            recv = send
        else:
            up, dn = (self.rank + 1) % self.world, (self.rank - 1) % self.world
            recv = torch.empty_like(send)
            for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, send, up, group), dist.P2POp(dist.irecv, recv, dn, group)]):
                req.wait()
        o = 0
        for a, b in runs:
            v[a:b, :, :, H - d:H] = recv[o:o + b - a]
            o += b - a

    def extend(self, interior, group=None):
        """[nb, V4_loc, 12] eigenvectors (local even/odd order) -> [nb, V4_ext, 12] with the halos of both neighbours.
        world == 1: the halos are the rank's own far slices (plain periodic lattice)."""
        return self.finish_extend(self.begin_extend(interior, group=group))

    def halo_bytes_per_vector(self, itemsize=16, slices=None):
        """bytes one eigenvector sends (= receives) per extension: `slices` time-slices (default: H to each of the two
        neighbours) x V3 sites x 12 complex"""
        return (2 * self.H if slices is None else slices) * 2 * self.V3h * 12 * itemsize

    def gather_time(self, local_mom, group=None):
        """[Nmom, nData, Tl] per rank -> [Nmom, nData, T] on every rank (COMM_TIME gather + broadcast of the reference)."""
        import torch.distributed as dist
        if self.world == 1:
            return local_mom
        parts = [torch.empty_like(local_mom) for _ in range(self.world)]
        dist.all_gather(parts, local_mom.contiguous(), group=group)
        return torch.cat(parts, dim=-1)
