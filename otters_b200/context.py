"""Device context: one CUDA device + one stream (``otters_ctx`` in include/otters_b200.h)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

from . import _ffi
from .types import OttersError


def check(rc: int) -> None:
    if rc != 0:
        raise OttersError(_ffi.last_error())


class Context:
    """Owns an ``otters_ctx``.  ``stream`` may be a raw ``cudaStream_t`` handle (e.g.
    ``torch.cuda.current_stream().cuda_stream``) so that work interleaves with the caller's."""

    def __init__(self, device: int = 0, stream: Optional[int] = None):
        h = C.c_void_p()
        if stream is not None and int(stream) == 0:
            raise OttersError("pass a real stream handle (the legacy default stream 0 is not supported) or None")
        check(_ffi.otters_ctx_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self._h = h
        self.device = int(device)
        self.stream = int(stream) if stream else None  # None: the library's own non-blocking stream

    @property
    def handle(self):
        return self._h

    def synchronize(self) -> None:
        check(_ffi.otters_ctx_synchronize(self._h))

    def join(self) -> None:
        """Orders this context's stream after all queries submitted on its lanes (no host wait)."""
        check(_ffi.otters_ctx_join(self._h))

    def set_tuning(self, warps_per_cta=0, slots_per_warp=0, kc_floats=0, ctas_per_sm=0, unit_rows=0, disable_fused_predicate=0,
                   batch_mode=0, batch_cta_group=0, scan_mode=0, planners=0, timing=0, batch_passes=0, separate_select=0,
                   lazy_prune=0) -> None:
        """batch_mode: 0 = automatic, 1 = always serve query batches with the tcgen05 kernel, 2 = never.
        batch_cta_group: 0 = automatic (single CTAs), 1 = single CTAs, 2 = CTA pairs (tcgen05 cta_group::2).
        scan_mode: K1 front-end, 0 = automatic, 1 = autonomous warps, 2 = planner + worker warps; planners: planner warps per CTA.
        timing: per-phase CUDA events; 0 = only for blocking MetaStore queries with stats, 1 = always, 2 = never.
        batch_passes: tensor-core kernel, 0 = bf16 selection first (bf16 shadow rows), then single-pass tf32, then 3xTF32, each when the
        certificate of the one before fails; 1 (tf32 single pass) / 2 (bf16) / 3 (3xTF32) = only that rung.
        separate_select: 1 = run the final selection (K3) as its own kernel instead of in the last CTA of the scan kernel.
        lazy_prune: 1 = evaluate the chunk rules (K0) per work unit inside the scan kernel instead of as their own kernel."""
        t = _ffi.ScanTuning(warps_per_cta, slots_per_warp, kc_floats, ctas_per_sm, unit_rows, disable_fused_predicate, batch_mode,
                            batch_cta_group, scan_mode, planners, timing, batch_passes, separate_select, lazy_prune)
        check(_ffi.otters_ctx_set_tuning(self._h, C.byref(t)))

    def last_work(self) -> Dict[str, float]:
        w = _ffi.LastWork()
        check(_ffi.otters_ctx_last_work(self._h, C.byref(w)))
        return {f: getattr(w, f) for f, _ in _ffi.LastWork._fields_}

    def close(self) -> None:
        if self._h:
            _ffi.otters_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default: Dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    """Process-wide context per device, created on first use (fails loudly without a B200)."""
    ctx = _default.get(device)
    if ctx is None:
        ctx = _default[device] = Context(device)
    return ctx
