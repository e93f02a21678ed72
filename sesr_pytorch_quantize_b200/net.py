"""Int8Net: the fused integer-sim network (one libsesrq handle), called with torch CUDA tensors.

PyTorch supplies device memory and streams only; all arithmetic happens in libsesrq's sm_100a
kernels.  This object is what the fused myQL graph (myQL/graph_modify.py mirror) executes, and what
bench.py and the data-parallel driver call.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from .compile import CompiledNet


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Int8Net:
    """Replaces GraphModule.forward of the sim.py graph (sim.py:82-114,205)."""

    def __init__(self, cnet: CompiledNet, device=None, impl: int = _lib.IMPL_AUTO, fuse: Optional[bool] = None):
        self.lib = _lib.load_library()
        if not torch.cuda.is_available():
            raise _lib.SesrqError("sesrq needs a CUDA device (B200 / sm_100); no CPU fallback exists")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.cnet = cnet
        self.impl = impl
        d = _lib.NetDesc()
        d.abi_version = _lib.ABI_VERSION
        d.n_layers = len(cnet.layers)
        d.r_shuffle = cnet.r_shuffle
        d.pe_num = cnet.pe_num
        d.pe_acc_bits = cnet.pe_acc_bits
        d.pe_add_bits = cnet.pe_add_bits
        d.quan_bits = cnet.quan_bits
        d.impl = impl
        d.scale_in = cnet.scales[0]
        d.scale_out = cnet.scales[5]
        d.z_out = cnet.zeros[5]
        d.res_m16 = cnet.res_m16
        d.res_shift_n = cnet.res_shift_n
        self._keep = []
        for i, L in enumerate(cnet.layers):
            w = L.weight.to("cpu", torch.int8).contiguous()
            self._keep.append(w)
            ld = d.layers[i]
            ld.ksize, ld.cin, ld.cout, ld.relu = L.ksize, L.cin, L.cout, int(L.relu)
            ld.z_in, ld.m16, ld.shift_n = L.z_in, L.m16, L.shift_n
            ld.h_weight = w.data_ptr()
            for j, v in enumerate(L.add_const):
                ld.add_const[j] = v
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.sesrq_net_create(C.byref(d), C.byref(h)), "sesrq_net_create")
        self._h = h
        if fuse is not None:            # None: the library default (on, unless SESRQ_FUSE=0)
            self.set_fusion(fuse)
        self._ws: Optional[torch.Tensor] = None
        self._ws_shape = None

    def close(self):
        if getattr(self, "_h", None):
            self.lib.sesrq_net_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection -----------------------------------------------------------------------
    def layer_info(self, layer: int) -> Tuple[int, int]:
        impl, pe = C.c_int32(), C.c_int32()
        _lib.check(self.lib.sesrq_net_layer_info(self._h, layer, C.byref(impl), C.byref(pe)))
        return impl.value, pe.value

    def layer_layout(self, layer: int) -> str:
        """Accumulator layout of a layer on the tcgen05 path: 'ns1' / 'spare' / 'ns2' / 'ns4', or 'generic' (dp4a path)."""
        v = self.lib.sesrq_net_layer_layout(self._h, layer)
        if v < 0:
            _lib.check(v, "sesrq_net_layer_layout")
        return ("generic", "ns1", "spare", "ns2", "ns4")[v]

    def set_fusion(self, enable: bool) -> bool:
        """Layers 1 + 2 in one launch (csrc/sesrq_chain.cu; bit-identical results, domain 2 is then never written to HBM,
        so read_activation(2) needs fusion off).  Returns the previous setting."""
        prev = self.lib.sesrq_net_set_fusion(self._h, int(bool(enable)))
        if prev < 0:
            _lib.check(prev, "sesrq_net_set_fusion")
        return bool(prev)

    @property
    def fused(self) -> bool:
        """True if the next forward runs layers 1 + 2 as one launch."""
        return bool(self.lib.sesrq_net_fused_layers(self._h))

    def ws_bytes(self, B: int, H: int, W: int) -> int:
        return int(self.lib.sesrq_ws_bytes(self._h, B, H, W))

    def _workspace(self, B, H, W) -> torch.Tensor:
        need = self.ws_bytes(B, H, W)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(max(need, 256), dtype=torch.uint8, device=self.device)
        self._ws_shape = (B, H, W)
        return self._ws

    # ---- forward ------------------------------------------------------------------------------
    def out_shapes(self, B, H, W):
        r = self.cnet.r_shuffle
        return (B, self.cnet.cout, H, W), (B, self.cnet.cout // (r * r), H * r, W * r)

    def forward(self, x: torch.Tensor, want_q5: bool = True, want_y: bool = True,
                q5: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None):
        """x: float32 CUDA NCHW.  Returns (q5 int8 [B,Cout,H,W] or None, y float32 pixel-shuffled or None)."""
        if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4:
            raise _lib.SesrqError("Int8Net.forward needs a float32 CUDA NCHW tensor")
        if x.shape[1] != self.cnet.cin:
            raise _lib.SesrqError("expected %d input channels, got %d" % (self.cnet.cin, x.shape[1]))
        x = x.contiguous()
        B, _, H, W = x.shape
        qs, ys = self.out_shapes(B, H, W)
        if want_q5 and q5 is None:
            q5 = torch.empty(qs, dtype=torch.int8, device=x.device)
        if want_y and y is None:
            y = torch.empty(ys, dtype=torch.float32, device=x.device)
        ws = self._workspace(B, H, W)
        with torch.cuda.device(x.device):
            rc = self.lib.sesrq_net_forward_i8(self._h, _ptr(x), B, H, W, _ptr(q5 if want_q5 else None),
                                               _ptr(y if want_y else None), _ptr(ws), ws.numel(),
                                               _stream_ptr(x.device))
        _lib.check(rc, "sesrq_net_forward_i8")
        return (q5 if want_q5 else None), (y if want_y else None)

    __call__ = forward

    def capture(self, x: torch.Tensor, want_q5: bool = False, want_y: bool = True) -> "GraphedInt8Net":
        """Capture the five launches of forward() for this input shape into a CUDA graph (small frames are bound by
        launch overhead: BASELINE config 1 is one 80x960 frame).  Returns a callable that copies its argument into the
        graph's static input and replays; outputs are the graph's static tensors."""
        return GraphedInt8Net(self, x, want_q5, want_y)

    def forward_host(self, x: torch.Tensor, q5: Optional[torch.Tensor] = None, y: Optional[torch.Tensor] = None):
        """Host-buffer entry point (sesrq_net_forward_i8_host): x, q5, y are CPU tensors (pinned
        memory recommended); H2D, the five layers and D2H all run inside the call."""
        if x.is_cuda or x.dtype != torch.float32:
            raise _lib.SesrqError("forward_host needs a float32 CPU tensor")
        x = x.contiguous()
        B, _, H, W = x.shape
        with torch.cuda.device(self.device):
            rc = self.lib.sesrq_net_forward_i8_host(self._h, _ptr(x), B, H, W, _ptr(q5), _ptr(y),
                                                    _stream_ptr(self.device))
        _lib.check(rc, "sesrq_net_forward_i8_host")
        return q5, y

    def profile(self, enable: bool):
        """Bracket every kernel of forward() with CUDA events (sesrq_net_profile_enable)."""
        _lib.check(self.lib.sesrq_net_profile_enable(self._h, int(enable)))

    def profile_read(self):
        """(ms[6], launches[6]) accumulated since the last read; index 5 = input-quantise pass."""
        ms = (C.c_double * (_lib.NUM_LAYERS + 1))()
        ln = (C.c_uint64 * (_lib.NUM_LAYERS + 1))()
        _lib.check(self.lib.sesrq_net_profile_read(self._h, ms, ln))
        return list(ms), list(ln)

    def read_activation(self, domain: int) -> torch.Tensor:
        """q_domain (input.{domain}.pt) of the last forward as int8 NCHW (layer-by-layer parity)."""
        B, H, W = self._ws_shape
        C_ = self.cnet.layers[0].cin if domain == 0 else self.cnet.layers[domain - 1].cout
        out = torch.empty((B, C_, H, W), dtype=torch.int8, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.sesrq_ws_read_act_i8(self._h, _ptr(self._ws), self._ws.numel(), B, H, W, domain,
                                               _ptr(out), _stream_ptr(self.device))
        _lib.check(rc, "sesrq_ws_read_act_i8")
        return out


class GraphedInt8Net:
    """Int8Net.forward for one fixed [B, C, H, W] as a replayable CUDA graph (tensor maps and kernel parameters are
    baked into the graph nodes; the workspace and the output tensors are static)."""

    def __init__(self, net: Int8Net, x: torch.Tensor, want_q5: bool, want_y: bool):
        self.net, self.want_q5, self.want_y = net, want_q5, want_y
        self.x = x.detach().clone().contiguous()
        B, _, H, W = self.x.shape
        qs, ys = net.out_shapes(B, H, W)
        self.q5 = torch.empty(qs, dtype=torch.int8, device=x.device) if want_q5 else None
        self.y = torch.empty(ys, dtype=torch.float32, device=x.device) if want_y else None
        net.profile(False)
        side = torch.cuda.Stream(x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):                      # warm-up off the capture: workspace, shared-memory opt-in
            for _ in range(2):
                net.forward(self.x, want_q5, want_y, q5=self.q5, y=self.y)
        torch.cuda.current_stream(x.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            net.forward(self.x, want_q5, want_y, q5=self.q5, y=self.y)

    def __call__(self, x: Optional[torch.Tensor] = None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.q5, self.y
