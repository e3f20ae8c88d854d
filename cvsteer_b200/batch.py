"""Device-resident batch + pyramid path over torch CUDA tensors.

torch is plumbing here: it owns device memory and streams; every kernel launched is ours, through the C ABI
(cvs_g2_run_batch_dev / cvs_g4_run_batch_dev / cvs_pyr_down_dev).  This is the reference's per-file loop
(example/steer.cpp:69-124) with the file I/O removed: frames in HBM in, selected planes in HBM out.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import capi


@dataclass
class Band:
    """Row band of one tall image: the input tensor holds image rows
    [y_origin, y_origin + tensor_rows) of an image `full_rows` tall; outputs are produced for image rows
    [row_begin, row_end) and land at row (y - row_begin) of the output tensors."""
    full_rows: int
    y_origin: int
    row_begin: int
    row_end: int


def _as_batch(x: torch.Tensor) -> torch.Tensor:
    if x.dim() == 2:
        x = x.unsqueeze(0)
    if x.dim() != 3 or not x.is_cuda or x.dtype not in (torch.float32, torch.uint8):
        raise capi.CvsError(capi.ERR_INVALID_ARG, "frames must be a CUDA tensor [n, rows, cols] of float32 or uint8")
    if x.stride(2) != 1:
        x = x.contiguous()
    return x


def _batch_struct(x: torch.Tensor, out_pitch: int, out_frame_stride: int, band: Optional[Band]) -> capi.Batch:
    b = capi.Batch()
    es = x.element_size()
    b.in_ = x.data_ptr()
    b.in_is_u8 = int(x.dtype == torch.uint8)
    b.n, b.rows, b.cols = x.shape
    b.in_pitch = x.stride(1) * es
    b.in_frame_stride = x.stride(0) * es if x.shape[0] > 1 else x.stride(1) * es * x.shape[1]
    b.out_pitch, b.out_frame_stride = out_pitch, out_frame_stride
    if band is not None:
        b.full_rows, b.y_origin = band.full_rows, band.y_origin
        b.out_row_begin, b.out_row_end, b.out_row_origin = band.row_begin, band.row_end, band.row_begin
    return b


def _stream_ptr(stream):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


class _FusedBatch:
    _prefix = ""
    _names: Sequence[str] = ()

    def __init__(self, width: int, spacing: float, device: Optional[int] = None):
        self._lib = capi.lib()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self._h = C.c_void_p()
        capi.check(getattr(self._lib, f"cvs_{self._prefix}_create")(C.byref(self._h), self.device, width, spacing))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            getattr(self._lib, f"cvs_{self._prefix}_destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, frames: torch.Tensor, mask: int, steer: int = capi.STEER_DOMINANT, theta: float = 0.0,
            theta_map: Optional[torch.Tensor] = None, outs: Optional[Dict[int, torch.Tensor]] = None,
            band: Optional[Band] = None, stream=None, next_level: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """One fused launch over the whole batch.  Returns {plane name: tensor [n, out_rows, cols]}.
        next_level: optional contiguous float32 [n, (rows+1)//2, (cols+1)//2]; the same launch fills it with
        cv::pyrDown(frames) from the tile it stages anyway (whole frames only)."""
        x = _as_batch(frames)
        if x.device.index != self.device:
            raise capi.CvsError(capi.ERR_INVALID_ARG, f"frames live on cuda:{x.device.index}, this handle on cuda:{self.device}")
        n, rows, cols = x.shape
        out_rows = rows if band is None else band.row_end - band.row_begin
        nplanes = len(self._names)
        planes = [p for p in range(nplanes) if mask >> p & 1]
        if outs is None:
            outs = {p: torch.empty((n, out_rows, cols), dtype=torch.float32, device=x.device) for p in planes}
        arr = (C.c_void_p * nplanes)()
        for p in planes:
            t = outs[p]
            if t.shape != (n, out_rows, cols) or t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
                raise capi.CvsError(capi.ERR_SIZE_MISMATCH, f"output plane {p}: need contiguous CUDA float32 {(n, out_rows, cols)}")
            arr[p] = t.data_ptr()
        tm = None
        if steer == capi.STEER_MAP:
            # the kernel reads theta_map + frame * out_frame_stride: the map needs the full [n, out_rows, cols] extent
            if theta_map is not None and theta_map.dim() == 2 and n == 1:
                theta_map = theta_map.unsqueeze(0)
            if (theta_map is None or tuple(theta_map.shape) != (n, out_rows, cols) or theta_map.dtype != torch.float32
                    or not theta_map.is_contiguous() or theta_map.device != x.device):
                raise capi.CvsError(capi.ERR_SIZE_MISMATCH,
                                    f"theta_map must be a contiguous float32 tensor {(n, out_rows, cols)} on {x.device}")
            tm = C.c_void_p(theta_map.data_ptr())
        b = _batch_struct(x, cols * 4, out_rows * cols * 4, band)
        if next_level is not None:
            want = (n, (rows + 1) // 2, (cols + 1) // 2)
            if tuple(next_level.shape) != want or next_level.dtype != torch.float32 or not next_level.is_contiguous():
                raise capi.CvsError(capi.ERR_SIZE_MISMATCH, f"next_level must be contiguous float32 {want}")
            b.next_level = next_level.data_ptr()
            b.next_pitch, b.next_frame_stride = want[2] * 4, want[1] * want[2] * 4
        fn = getattr(self._lib, f"cvs_{self._prefix}_run_batch_dev")
        with torch.cuda.device(self.device):   # the C call makes the handle's GPU current; torch's current device is restored
            capi.check(fn(self._h, C.byref(b), mask, steer, theta, tm, arr, _stream_ptr(stream)))
        return {self._names[p]: outs[p] for p in planes}

    def run_pyramid(self, frames: torch.Tensor, levels: int, mask: int, **kw) -> List[Dict[str, torch.Tensor]]:
        """Every pyramid level stays resident on the device: level l+1 = pyr_down(level l), then one fused
        launch per level (config 3 of BASELINE.json)."""
        res, cur = [], _as_batch(frames)
        fuse = kw.pop("fuse_pyramid", True)
        for l in range(levels):
            nxt = None
            if l + 1 < levels and fuse:
                n, r, c = cur.shape
                nxt = torch.empty((n, (r + 1) // 2, (c + 1) // 2), dtype=torch.float32, device=cur.device)
            res.append(self.run(cur, mask, next_level=nxt, **kw))
            if l + 1 < levels:
                cur = nxt if fuse else pyr_down(cur, stream=kw.get("stream"))
        return res

    def last_launch(self):
        g = (C.c_int * 3)()
        blk, sm = C.c_int(), C.c_int()
        name = C.create_string_buffer(96)
        capi.check(self._lib.cvs_g2_last_launch(self._h, g, C.byref(blk), C.byref(sm), name, 96))
        return {"grid": list(g), "block": blk.value, "smem": sm.value, "kernel": name.value.decode()}


class G2Batch(_FusedBatch):
    _prefix = "g2"
    _names = capi.G2_PLANE_NAMES

    def __init__(self, width: int = 4, spacing: float = 0.67, device: Optional[int] = None):
        super().__init__(width, spacing, device)

    def run_host(self, frames_host: torch.Tensor, mask: int, outs_host: Dict[int, torch.Tensor]):
        """End-to-end host-buffer call (cvs_g2_run_batch_host): pinned host frames in, pinned host planes out."""
        x = frames_host
        n, rows, cols = x.shape
        arr = (C.c_void_p * capi.G2_NPLANES)()
        for p, t in outs_host.items():
            arr[p] = t.data_ptr()
        any_out = next(iter(outs_host.values()))
        capi.check(self._lib.cvs_g2_run_batch_host(self._h, C.c_void_p(x.data_ptr()), n, rows, cols, x.stride(1) * 4,
                                                   x.stride(0) * 4, mask, arr, any_out.stride(1) * 4,
                                                   any_out.stride(0) * 4))


class G4Batch(_FusedBatch):
    _prefix = "g4"
    _names = capi.G4_PLANE_NAMES

    def __init__(self, width: int = 6, spacing: float = 0.5, device: Optional[int] = None):
        super().__init__(width, spacing, device)

    def run_host(self, frames_host: torch.Tensor, mask: int, outs_host: Dict[int, torch.Tensor],
                 theta: Optional[float] = None):
        """End-to-end host-buffer call (cvs_g4_run_batch_host).  theta=None steers at the G4 theta_d."""
        x = frames_host
        n, rows, cols = x.shape
        arr = (C.c_void_p * capi.G4_NPLANES)()
        for p, t in outs_host.items():
            arr[p] = t.data_ptr()
        any_out = next(iter(outs_host.values()))
        src = capi.STEER_DOMINANT if theta is None else capi.STEER_SCALAR
        capi.check(self._lib.cvs_g4_run_batch_host(self._h, C.c_void_p(x.data_ptr()), n, rows, cols, x.stride(1) * 4,
                                                   x.stride(0) * 4, mask, src, float(theta or 0.0), arr,
                                                   any_out.stride(1) * 4, any_out.stride(0) * 4))


def pyr_down(frames: torch.Tensor, band: Optional[Band] = None, out_rows: Optional[range] = None,
             stream=None) -> torch.Tensor:
    """cv::pyrDown semantics on a device batch.  With `band`, `frames` holds rows [y_origin, ...) of the input
    level and `band.row_begin/row_end` are rows of the OUTPUT level."""
    x = _as_batch(frames)
    n, rows, cols = x.shape
    oc = (cols + 1) // 2
    orows = (rows + 1) // 2 if band is None else band.row_end - band.row_begin
    out = torch.empty((n, orows, oc), dtype=torch.float32, device=x.device)
    b = _batch_struct(x, oc * 4, orows * oc * 4, band)
    capi.check(capi.lib().cvs_pyr_down_dev(x.device.index or 0, C.byref(b), C.c_void_p(out.data_ptr()),
                                           _stream_ptr(stream)))
    return out


def ffma_peak(form: int = 2, iters: int = 20000, device: int = 0):
    """Measured FP32 FFMA issue rate (instr/s) for the roofline denominator."""
    v, ms = C.c_double(), C.c_float()
    capi.check(capi.lib().cvs_bench_ffma(device, form, iters, C.byref(v), C.byref(ms)))
    return v.value, ms.value
