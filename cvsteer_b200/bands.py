"""Row-band runs of ONE large image over several GPUs through the C ABI (cvs_bands_*, include/cvsteer_c.h).

One process per GPU.  torch.distributed is only the CONTROL plane here (it carries the root's 64-byte CUDA IPC handle and
the 128-byte NCCL unique id to the other ranks); everything on the data path -- kernels, peer stores, copy-engine
transfers, ncclSend/ncclRecv -- is issued by libcvsteer_b200 itself."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import capi

GATHER = {"none": capi.GATHER_NONE, "nccl": capi.GATHER_NCCL, "peer": capi.GATHER_PEER_STORE, "peer_store": capi.GATHER_PEER_STORE,
          "copy": capi.GATHER_PEER_COPY, "peer_copy": capi.GATHER_PEER_COPY}


class _RawCuda:
    """A device pointer dressed as a __cuda_array_interface__ exporter so that torch can alias it without copying."""

    def __init__(self, ptr: int, nelem: int):
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def _view(ptr: int, rows: int, cols: int, pitch: int, device: int) -> torch.Tensor:
    if rows <= 0:
        return torch.empty((0, cols), dtype=torch.float32, device=torch.device("cuda", device))
    flat = torch.as_tensor(_RawCuda(ptr, rows * pitch // 4), device=torch.device("cuda", device))
    return flat.view(rows, pitch // 4)[:, :cols]


class BandRun:
    """One rank of a row-band run.  With world > 1 the constructor is COLLECTIVE over torch.distributed's default group
    (or `group`): it exchanges the root's IPC handle and, unless nccl=False, creates the library's own NCCL communicator."""

    def __init__(self, rows: int, cols: int, levels: int, device: Optional[int] = None, world: int = 1, rank: int = 0,
                 root: int = 0, mask: int = capi.G2_MASK_ORIENT, width: int = 4, spacing: float = 0.67, group=None,
                 nccl: bool = True, peer: bool = True):
        self._lib = capi.lib()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.rows, self.cols, self.levels, self.world, self.rank, self.root, self.mask = rows, cols, levels, world, rank, root, mask
        self.group = group
        self._h = C.c_void_p()
        capi.check(self._lib.cvs_bands_create(C.byref(self._h), self.device, rank, world, root, rows, cols, levels, mask, width, spacing))
        self.has_root_planes = False
        self.has_nccl = False
        if world > 1:
            import torch.distributed as dist
            if peer:
                obj = [None]
                if rank == root:
                    handle = C.create_string_buffer(64)
                    capi.check(self._lib.cvs_bands_root_export(self._h, handle, None))
                    obj = [handle.raw]
                dist.broadcast_object_list(obj, src=root, group=group)
                if rank != root:
                    capi.check(self._lib.cvs_bands_root_import(self._h, obj[0]))
                self.has_root_planes = True
            elif rank == root:
                capi.check(self._lib.cvs_bands_root_export(self._h, None, None))
                self.has_root_planes = True
            if nccl:
                obj = [None]
                if rank == 0:
                    uid = C.create_string_buffer(128)
                    capi.check(self._lib.cvs_nccl_unique_id(uid))
                    obj = [uid.raw]
                dist.broadcast_object_list(obj, src=0, group=group)
                with torch.cuda.device(self.device):
                    capi.check(self._lib.cvs_bands_nccl_init(self._h, obj[0]))
                self.has_nccl = True
            dist.barrier(group)
        else:
            capi.check(self._lib.cvs_bands_root_export(self._h, None, None))
            self.has_root_planes = True

    # ---- geometry ----
    def geometry(self, level: int, rank: int = -1) -> Dict[str, int]:
        v = [C.c_int() for _ in range(6)]
        pitch = C.c_size_t()
        capi.check(self._lib.cvs_bands_geometry(self._h, rank, level, *[C.byref(x) for x in v], C.byref(pitch)))
        keys = ("level_rows", "level_cols", "out_lo", "out_hi", "have_lo", "have_hi")
        d = {k: x.value for k, x in zip(keys, v)}
        d["pitch"] = pitch.value
        return d

    def band_rows(self) -> int:
        g = self.geometry(0)
        return g["out_hi"] - g["out_lo"]

    def halo_rows(self) -> int:
        g = self.geometry(0)
        return (g["have_hi"] - g["have_lo"]) - (g["out_hi"] - g["out_lo"])

    def gather_bytes_to_root(self) -> int:
        """Payload bytes that cross NVLink into the root per step (every other rank's rows of every selected plane)."""
        nsel = bin(self.mask).count("1")
        total = 0
        for r in range(self.world):
            if r == self.root:
                continue
            for l in range(self.levels):
                g = self.geometry(l, r)
                total += max(0, g["out_hi"] - g["out_lo"]) * g["level_cols"] * 4 * nsel
        return total

    # ---- input ----
    def input(self) -> torch.Tensor:
        """This rank's level-0 buffer: image rows [have_lo, have_hi) as a [rows, cols] view (row stride = pitch)."""
        p, pitch = C.c_void_p(), C.c_size_t()
        capi.check(self._lib.cvs_bands_input_dev(self._h, C.byref(p), C.byref(pitch)))
        g = self.geometry(0)
        return _view(p.value or 0, g["have_hi"] - g["have_lo"], self.cols, pitch.value, self.device)

    def load_synthetic(self, seed: int = 0):
        x = self.input()
        if x.numel():
            gen = torch.Generator(device=x.device)
            gen.manual_seed(seed + self.rank)
            x.copy_(torch.rand(x.shape, device=x.device, generator=gen) * 255.0)

    def load_image_rows(self, image: torch.Tensor):
        """Copy this rank's rows out of a whole image (host or device tensor [rows, cols])."""
        g = self.geometry(0)
        x = self.input()
        if x.numel():
            x.copy_(image[g["have_lo"]:g["have_hi"]].to(x.device, non_blocking=False))

    # ---- run ----
    def step(self, gather: str = "peer", stream=None, barrier: bool = True):
        s = stream if stream is not None else torch.cuda.current_stream()
        mode = GATHER[gather]
        with torch.cuda.device(self.device):
            capi.check(self._lib.cvs_bands_run(self._h, mode, C.c_void_p(s.cuda_stream)))
            if barrier and self.world > 1 and mode in (capi.GATHER_PEER_STORE, capi.GATHER_PEER_COPY) and self.has_nccl:
                capi.check(self._lib.cvs_bands_barrier(self._h, C.c_void_p(s.cuda_stream)))

    # ---- results ----
    def _planes(self, fn, level: int, rows: int) -> Dict[str, torch.Tensor]:
        out = {}
        cols = self.geometry(level)["level_cols"]
        for p in range(capi.G2_NPLANES):
            if self.mask >> p & 1:
                ptr, pitch = C.c_void_p(), C.c_size_t()
                capi.check(fn(self._h, level, p, C.byref(ptr), C.byref(pitch)))
                out[capi.G2_PLANE_NAMES[p]] = _view(ptr.value or 0, rows, cols, pitch.value, self.device)
        return out

    def root_planes(self, level: int) -> Dict[str, torch.Tensor]:
        """Full-size planes of `level` (valid on the root; a peer mapping of the same memory elsewhere)."""
        return self._planes(self._lib.cvs_bands_root_plane, level, self.geometry(level)["level_rows"])

    def local_planes(self, level: int) -> Dict[str, torch.Tensor]:
        g = self.geometry(level)
        return self._planes(self._lib.cvs_bands_local_plane, level, g["out_hi"] - g["out_lo"])

    def close(self):
        """Collective with world > 1: every rank detaches at the same time (NCCL communicator, imported mapping), then a
        barrier, then the contexts are destroyed (the root frees its block after its peers have unmapped it)."""
        if getattr(self, "_h", None) and self._h.value:
            torch.cuda.synchronize()
            if self.world > 1:
                import torch.distributed as dist
                with torch.cuda.device(self.device):
                    capi.check(self._lib.cvs_bands_detach(self._h))
                dist.barrier(self.group)
            self._lib.cvs_bands_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                self._lib.cvs_bands_destroy(self._h)
        except Exception:
            pass
