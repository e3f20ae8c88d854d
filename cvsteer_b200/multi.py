"""Multi-GPU dispatch: one process per GPU (torch.distributed), two sharding modes.

* frames  -- independent frames split contiguously over ranks; NO collective on the data path (the reference's own
             parallelism is per file: example/steer.cpp:169).
* bands   -- one very large image split into row bands.  Band edges sit on multiples of 2^(levels-1) rows so that the
             even-sample pyramid of a band coincides with the pyramid of the whole image; every rank loads its band
             PLUS the halo rows its coarsest level needs from the (host-resident) input, so no halo exchange is needed;
             the only communication is ONE gather of the output bands to a root (NCCL send/recv over NVLink, posted
             as a single batch so all peers stream concurrently).  Results are bit-identical to the whole-image run.

             Alternative to the NCCL gather: `PeerPlanes` maps the root's full-size output planes into every rank
             (CUDA IPC over NVLink peer memory), and each rank's fused kernel STORES its rows straight into them --
             compute and gather are one kernel, the transfer overlaps the arithmetic tile by tile, and no staging copy
             or second pass over the outputs exists on any rank.

The planning functions are pure Python (tested on CPU); `process` callables keep the compute injectable so that the
world_size-2 gloo tests exercise exactly the code path the NCCL run takes.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

G2_RADIUS = 4   # default G2/H2 width
PYR_RADIUS = 2  # [1 4 6 4 1]


# ------------------------------------------------------------------------------------------------
# frames
# ------------------------------------------------------------------------------------------------
def shard_frames(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of ceil(n/world) frames per rank (trailing ranks may get fewer or none)."""
    per = -(-n_frames // world)
    lo = min(n_frames, rank * per)
    return lo, min(n_frames, lo + per)


# ------------------------------------------------------------------------------------------------
# bands
# ------------------------------------------------------------------------------------------------
def level_rows(rows: int, levels: int) -> List[int]:
    out = [rows]
    for _ in range(levels - 1):
        out.append((out[-1] + 1) // 2)
    return out


@dataclass
class BandPlan:
    """Per-rank row bookkeeping for every pyramid level (all ranges are [lo, hi) in that level's image rows)."""
    rank: int
    rows: List[int]                                           # image height per level
    out: List[Tuple[int, int]] = field(default_factory=list)   # rows this rank PRODUCES per level
    have: List[Tuple[int, int]] = field(default_factory=list)  # rows of each level this rank must hold (out + halos)

    @property
    def empty(self) -> bool:
        return self.out[0][0] >= self.out[0][1]


def plan_bands(rows: int, world: int, levels: int, radius: int = G2_RADIUS) -> List[BandPlan]:
    align = 1 << (levels - 1)
    hl = level_rows(rows, levels)
    units = -(-rows // align)
    per = -(-units // world)
    edges = [min(rows, r * per * align) for r in range(world + 1)]
    edges[world] = rows
    plans = []
    for r in range(world):
        p = BandPlan(rank=r, rows=hl)
        for l in range(levels):
            if edges[r] >= rows:                      # nothing left for this rank (more ranks than aligned bands)
                p.out.append((hl[l], hl[l]))
                continue
            lo = min(hl[l], edges[r] >> l)
            hi = hl[l] if edges[r + 1] >= rows else min(hl[l], edges[r + 1] >> l)   # the last band runs to the end
            p.out.append((lo, max(lo, hi)))
        # rows each level must hold: what the basis kernel reads, plus what pyr_down of the next level reads
        have = [None] * levels
        for l in range(levels - 1, -1, -1):
            lo, hi = p.out[l]
            if lo >= hi:
                need = (lo, lo)
            else:
                need = (max(0, lo - radius), min(hl[l], hi + radius))
            if l + 1 < levels and have[l + 1][0] < have[l + 1][1]:
                a, b = have[l + 1]
                src = (max(0, 2 * a - PYR_RADIUS), min(hl[l], 2 * (b - 1) + PYR_RADIUS + 1))
                need = (min(need[0], src[0]), max(need[1], src[1])) if need[0] < need[1] else src
            have[l] = need
        p.have = have
        plans.append(p)
    return plans


ProcessBand = Callable[[torch.Tensor, int, BandPlan], Dict[str, torch.Tensor]]
DownBand = Callable[[torch.Tensor, int, BandPlan], torch.Tensor]


def plane_layout(names: Sequence[str], rows_per_level: Sequence[int], cols: int) -> Tuple[Dict[Tuple[int, str], Tuple[int, int, int]], int]:
    """Element offsets of every (level, plane) inside ONE flat fp32 buffer: {(level, name): (offset, rows, cols)}, total.
    Each plane starts on a 128-byte boundary (32 floats) so that TMA-describable consumers can use it as is."""
    out, off, c = {}, 0, cols
    for l, r in enumerate(rows_per_level):
        for n in names:
            out[(l, n)] = (off, r, c)
            off += -(-(r * c) // 32) * 32
        c = (c + 1) // 2
    return out, off


class _RawCuda:
    """A device pointer dressed as a __cuda_array_interface__ exporter so that torch can alias it without copying."""

    def __init__(self, ptr: int, nelem: int):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class PeerPlanes:
    """Root-owned full-size output planes of a band run, mapped into every rank over CUDA IPC (NVLink peer memory).

    full[level][name] is a [rows_l, cols_l] fp32 tensor on every rank: real memory on `root`, a peer mapping of the
    same memory elsewhere.  The block is allocated and exported by the C library (cvs_shared_alloc) and opened by
    every other rank with ITS OWN GPU current (cvs_shared_open) -- that is what makes the mapping usable by that
    GPU's kernels.  Create once per (image size, plane set) and reuse across steps; call close() (collective) when
    done.  Ranks call `fence()` after their kernels before the root reads the planes."""

    def __init__(self, names: Sequence[str], rows: int, cols: int, levels: int, root: int = 0, group=None):
        import ctypes as C

        from . import capi
        self.root, self.group = root, group
        self.rank = dist.get_rank(group)
        self.device = torch.cuda.current_device()
        self.names = list(names)
        self.layout, total = plane_layout(self.names, level_rows(rows, levels), cols)
        lib, ptr = capi.lib(), C.c_void_p()
        obj = [None]
        if self.rank == root:
            handle = C.create_string_buffer(64)
            capi.check(lib.cvs_shared_alloc(self.device, total * 4, C.byref(ptr), handle))
            obj = [handle.raw]
        dist.broadcast_object_list(obj, src=root, group=group)
        if self.rank != root:
            capi.check(lib.cvs_shared_open(self.device, obj[0], C.byref(ptr)))
        self._ptr = ptr.value
        self._flat = torch.as_tensor(_RawCuda(self._ptr, total), device=torch.device("cuda", self.device))
        self.full: List[Dict[str, torch.Tensor]] = [dict() for _ in range(levels)]
        for (l, n), (off, r, c) in self.layout.items():
            self.full[l][n] = self._flat[off:off + r * c].view(r, c)
        dist.barrier(group)               # nobody proceeds before every rank holds its mapping

    def rows_of(self, level: int, lo: int, hi: int) -> Dict[str, torch.Tensor]:
        """Views of rows [lo, hi) of every plane of `level` (contiguous row blocks)."""
        return {n: t[lo:hi] for n, t in self.full[level].items()}

    def fence(self):
        """All ranks' stores have landed in the root's memory when this returns on the root."""
        torch.cuda.current_stream().synchronize()   # kernel completion flushes its peer stores (system-scope release)
        dist.barrier(self.group)

    def close(self):
        """Collective: importers unmap, then the owner frees.  Tensors handed out by `full` are invalid afterwards."""
        from . import capi
        if self._ptr is None:
            return
        torch.cuda.synchronize()
        self.full, self._flat = [], None
        if self.rank != self.root:
            capi.check(capi.lib().cvs_shared_close(self.device, self._ptr))
        dist.barrier(self.group)
        if self.rank == self.root:
            capi.check(capi.lib().cvs_shared_free(self.device, self._ptr))
        self._ptr = None


def run_bands(load_rows: Callable[[int, int], torch.Tensor], rows: int, cols: int, levels: int, process: ProcessBand,
              down: DownBand, root: int = 0, gather=True, radius: int = G2_RADIUS,
              group=None, peer: Optional[PeerPlanes] = None) -> Tuple[Optional[List[Dict[str, torch.Tensor]]], List[Dict[str, torch.Tensor]], BandPlan]:
    """Row-band pipeline on the calling rank.

    load_rows(lo, hi) -> device tensor [hi-lo, cols] with image rows [lo, hi) of level 0 (band + halo).
    process(buf, level, plan) -> {name: tensor [out_rows_l, cols_l]} for rows plan.out[level], given `buf` holding
        rows plan.have[level] of that level.
    down(buf, level, plan) -> tensor holding rows plan.have[level+1] of level+1 (pyr_down in band mode).
    Returns (gathered per-level dicts on root or None, local per-level dicts, plan).

    gather: True / "nccl" -- compute into local planes, then ONE batched NCCL send/recv to the root;
            "direct"      -- needs `peer` (PeerPlanes): `process(buf, level, plan, outs=...)` writes this rank's rows
                             straight into the root's planes over NVLink; the only synchronisation is peer.fence();
            False         -- no gather.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    plans = plan_bands(rows, world, levels, radius)
    plan = plans[rank]
    local: List[Dict[str, torch.Tensor]] = []
    direct = gather == "direct"
    if direct and peer is None:
        raise ValueError("gather='direct' needs a PeerPlanes")
    if not plan.empty:
        buf = load_rows(*plan.have[0])
        for l in range(levels):
            if direct:
                local.append(process(buf, l, plan, outs=peer.rows_of(l, *plan.out[l])))
            else:
                local.append(process(buf, l, plan))
            if l + 1 < levels:
                buf = down(buf, l, plan)
    if direct:
        peer.fence()
        return (peer.full if rank == root else None), local, plan
    if not gather or world == 1:
        return (local if rank == root else None), local, plan
    return gather_bands(local, plans, cols, root, group), local, plan


def gather_bands(local: List[Dict[str, torch.Tensor]], plans: Sequence[BandPlan], cols: int, root: int = 0,
                 group=None) -> Optional[List[Dict[str, torch.Tensor]]]:
    """ONE batched exchange: every rank sends each of its (level, plane) row blocks to `root`, which receives them
    straight into row slices of the full-size outputs (row blocks are contiguous, so there is no staging copy)."""
    rank = dist.get_rank(group)
    levels = len(plans[0].rows)
    # plane names/dtypes must agree across ranks; take them from the first non-empty rank's result on the root side
    names = sorted(local[0].keys()) if local else None
    obj = [names]
    src_rank = next(p.rank for p in plans if not p.empty)
    dist.broadcast_object_list(obj, src=src_rank, group=group)
    names = obj[0]
    ops, full = [], None
    lc = [cols]
    for _ in range(levels - 1):
        lc.append((lc[-1] + 1) // 2)
    if rank == root:
        ref = next(iter(local[0].values())) if local else None
        dev = ref.device if ref is not None else torch.device("cuda", torch.cuda.current_device())
        full = [{n: torch.empty((plans[0].rows[l], lc[l]), dtype=torch.float32, device=dev) for n in names} for l in range(levels)]
        for p in plans:
            if p.empty:
                continue
            for l in range(levels):
                lo, hi = p.out[l]
                if lo >= hi:
                    continue
                for n in names:
                    dst = full[l][n][lo:hi]
                    if p.rank == root:
                        dst.copy_(local[l][n].reshape(hi - lo, lc[l]))
                    else:
                        ops.append(dist.P2POp(dist.irecv, dst, p.rank, group))
    else:
        p = plans[rank]
        if not p.empty:
            for l in range(levels):
                lo, hi = p.out[l]
                if lo >= hi:
                    continue
                for n in names:
                    ops.append(dist.P2POp(dist.isend, local[l][n].reshape(hi - lo, lc[l]).contiguous(), root, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return full


# ------------------------------------------------------------------------------------------------
# default CUDA callables (G2/H2 through the C ABI)
# ------------------------------------------------------------------------------------------------
def cuda_callables(mask: int, width: int = 4, spacing: float = 0.67):
    """process/down callables for run_bands that launch the fused kernels in band mode."""
    from .batch import Band, G2Batch, pyr_down

    g2 = G2Batch(width, spacing)

    from . import capi
    ids = {capi.G2_PLANE_NAMES[p]: p for p in range(capi.G2_NPLANES) if mask >> p & 1}

    def process(buf: torch.Tensor, level: int, plan: BandPlan, outs: Optional[Dict[str, torch.Tensor]] = None):
        lo, hi = plan.out[level]
        band = Band(full_rows=plan.rows[level], y_origin=plan.have[level][0], row_begin=lo, row_end=hi)
        if outs is not None:   # caller-owned destinations (possibly peer memory of another GPU): [rows, cols] row blocks
            outs = {ids[k]: v.unsqueeze(0) for k, v in outs.items()}
        res = g2.run(buf, mask, band=band, outs=outs)
        return {k: v[0] for k, v in res.items()}

    def down(buf: torch.Tensor, level: int, plan: BandPlan):
        a, b = plan.have[level + 1]
        band = Band(full_rows=plan.rows[level], y_origin=plan.have[level][0], row_begin=a, row_end=b)
        return pyr_down(buf, band=band)[0]

    return process, down, g2
