"""quantized-cnn_b200: Python binding (ctypes) of the B200-native PQ forward path.

The product is ``libqcnn_b200.so`` (hand-written sm_100a CUDA + C++ host code behind the C ABI of
``include/qcnn.h``).  This module only marshals pointers: torch supplies device memory and streams, every
computation happens inside the shared library.  There is no CPU or PyTorch fallback -- if the library is
missing or no B200 is visible the calls raise.

The directory name contains a hyphen, so import it with::

    import importlib; q = importlib.import_module("quantized-cnn_b200")
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqcnn_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "qcnn.h")


class QcnnError(RuntimeError):
    pass


def build(verbose=False):
    """Compile libqcnn_b200.so in-tree with nvcc for sm_100a (see Makefile)."""
    import subprocess
    cmd = ["make", "-C", _HERE, "-j8"]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)


def _load():
    if not os.path.exists(LIB_PATH):
        raise QcnnError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no fallback path)" % LIB_PATH)
    return C.CDLL(LIB_PATH)


lib = _load()


class LayerInfo(C.Structure):
    """== qcnn_layer_info == the reference's LayerInfo (include/CaffePara.h:28-41)."""
    _fields_ = [("type", C.c_int), ("padSiz", C.c_int), ("knlSiz", C.c_int), ("knlCnt", C.c_int),
                ("grpCnt", C.c_int), ("stride", C.c_int), ("nodCnt", C.c_int), ("lrnSiz", C.c_int),
                ("lrnAlp", C.c_float), ("lrnBet", C.c_float), ("lrnIni", C.c_float), ("drpRat", C.c_float)]


_vp, _i, _f, _sz, _cp = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_char_p
_dp = C.POINTER(C.c_double)
_SIGS = {
    "qcnn_version": (_cp, []),
    "qcnn_last_error": (_cp, []),
    "qcnn_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "qcnn_ctx_destroy": (None, [_vp]),
    "qcnn_ctx_device": (_i, [_vp]),
    "qcnn_ctx_sm_count": (_i, [_vp]),
    "qcnn_dev_alloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "qcnn_dev_free": (_i, [_vp, _vp]),
    "qcnn_copy_h2d": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "qcnn_copy_d2h": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "qcnn_stream_sync": (_i, [_vp, _vp]),
    "qcnn_conv_layer_create": (_i, [_vp] + [_i] * 11 + [_vp, _vp, _vp, C.POINTER(_vp)]),
    "qcnn_fc_layer_create": (_i, [_vp] + [_i] * 5 + [_vp, _vp, _vp, C.POINTER(_vp)]),
    "qcnn_fc_layer_set_src_nhwc": (_i, [_vp, _i, _i, _i]),
    "qcnn_conv_layer_set_src_nchw": (_i, [_vp, _i]),
    "qcnn_layer_set_param": (_i, [_vp, _cp, _i]),
    "qcnn_layer_describe": (_i, [_vp, _i, _cp, _sz]),
    "qcnn_layer_destroy": (None, [_vp]),
    "qcnn_layer_out_dims": (_i, [_vp, C.POINTER(_i)]),
    "qcnn_layer_work": (_i, [_vp, _i, _dp, _dp, _dp]),
    "qcnn_layer_read_asmt_h": (_i, [_vp, _vp, _sz]),
    "qcnn_conv_aprx_forward": (_i, [_vp, _vp, _i, _vp, _i, _vp]),
    "qcnn_fc_aprx_forward": (_i, [_vp, _vp, _i, _vp, _i, _vp]),
    "qcnn_fc_aprx_forward_flat": (_i, [_vp, _vp, _i, _vp, _i, _vp]),
    "qcnn_fc_chain_forward": (_i, [C.POINTER(_vp), C.POINTER(_i), _i, _vp, _i, _vp, _vp, _vp]),
    "qcnn_relu": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "qcnn_lrn": (_i, [_vp, _vp, _vp, _sz, _i, _i, _f, _f, _f, _vp]),
    "qcnn_maxpool": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "qcnn_lrn_maxpool": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _f, _f, _i, _i, _i, _vp]),
    "qcnn_softmax": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "qcnn_nchw_to_nhwc": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "qcnn_nhwc_to_nchw": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "qcnn_net_create": (_i, [_vp, _cp, _cp, _cp, C.POINTER(_vp)]),
    "qcnn_net_create_custom": (_i, [_vp, _i, C.POINTER(LayerInfo), _i, _i, _i, _cp, _cp, C.POINTER(_vp)]),
    "qcnn_net_create_from_para": (_i, [_vp, _i, C.POINTER(LayerInfo), _vp, _i, _i, _i, C.POINTER(_vp)]),
    "qcnn_net_destroy": (None, [_vp]),
    "qcnn_net_layer_count": (_i, [_vp]),
    "qcnn_net_out_len": (_i, [_vp]),
    "qcnn_net_set_keep_maps": (_i, [_vp, _i]),
    "qcnn_net_forward": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "qcnn_net_forward_h": (_i, [_vp, _vp, _i, _vp, _vp]),
    "qcnn_net_forward_topk_h": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "qcnn_net_set_input_mean": (_i, [_vp, _vp]),
    "qcnn_net_forward_u8": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "qcnn_net_forward_u8_h": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "qcnn_net_submit_u8_h": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, C.POINTER(_i)]),
    "qcnn_net_wait": (_i, [_vp, _i]),
    "qcnn_net_set_chunk": (_i, [_vp, _i]),
    "qcnn_net_featmap": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_i)]),
    "qcnn_net_set_profiling": (_i, [_vp, _i]),
    "qcnn_net_layer_time_ms": (_i, [_vp, _i, C.POINTER(_f)]),
    "qcnn_net_layer_work": (_i, [_vp, _i, _i, _dp, _dp, _dp]),
    "qcnn_net_launch_count": (_i, [_vp]),
    "qcnn_net_pq_layer": (_vp, [_vp, _i]),
    "qcnn_topk": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "qcnn_u8hwc_to_f32chw": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "qcnn_preproc_create": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _i, C.POINTER(_vp)]),
    "qcnn_preproc_destroy": (None, [_vp]),
    "qcnn_preproc_run": (_i, [_vp, _vp, C.POINTER(C.c_longlong), C.POINTER(_i), C.POINTER(_i), _i, _vp, _vp]),
    "qcnn_multi_create": (_i, [_i, C.POINTER(_i), _cp, _cp, _cp, C.POINTER(_vp)]),
    "qcnn_multi_create_from_para": (_i, [_i, C.POINTER(_i), _i, C.POINTER(LayerInfo), _vp, _i, _i, _i, C.POINTER(_vp)]),
    "qcnn_multi_destroy": (None, [_vp]),
    "qcnn_multi_device_count": (_i, [_vp]),
    "qcnn_multi_out_len": (_i, [_vp]),
    "qcnn_multi_nccl_version": (_i, [_vp]),
    "qcnn_multi_net": (_vp, [_vp, _i]),
    "qcnn_multi_forward_h": (_i, [_vp, _vp, _i, _vp]),
    "qcnn_multi_forward": (_i, [_vp, C.POINTER(_vp), _i, C.POINTER(_vp)]),
    "qcnn_multi_sync": (_i, [_vp]),
    "qcnn_read_bin_f32": (C.c_long, [_cp, C.POINTER(_i), C.POINTER(_i), _vp, C.c_long]),
    "qcnn_write_bin_f32": (_i, [_cp, _i, C.POINTER(_i), _vp]),
    "qcnn_read_cbn_u8": (C.c_long, [_cp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), _vp, C.c_long]),
    "qcnn_write_cbn_u8": (_i, [_cp, _i, C.POINTER(_i), _vp, _i]),
}
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args
EXPORTS = sorted(_SIGS)


def last_error():
    return lib.qcnn_last_error().decode()


def _check(rc):
    if rc != 0:
        raise QcnnError(last_error())


def _np(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _stream(stream):
    if stream is not None:
        return C.c_void_p(int(stream))
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dptr(t):
    """Device pointer of a contiguous float32 CUDA tensor."""
    import torch
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise QcnnError("expected a contiguous float32 CUDA tensor")
    return C.c_void_p(t.data_ptr())


class Context(object):
    def __init__(self, device=0):
        h = _vp()
        _check(lib.qcnn_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    @property
    def sm_count(self):
        return lib.qcnn_ctx_sm_count(self.h)

    def close(self):
        if self.h:
            lib.qcnn_ctx_destroy(self.h)
            self.h = None

    # ---- supporting layers; tensors are NHWC float32 on the context's device ----
    def relu(self, x, stream=None):
        import torch
        y = torch.empty_like(x)
        _check(lib.qcnn_relu(self.h, _dptr(x), _dptr(y), x.numel(), _stream(stream)))
        return y

    def lrn(self, x, size, alpha, beta, k, stream=None):
        import torch
        y = torch.empty_like(x)
        _check(lib.qcnn_lrn(self.h, _dptr(x), _dptr(y), x.numel() // x.shape[-1], x.shape[-1], size, alpha, beta, k,
                            _stream(stream)))
        return y

    @staticmethod
    def _pool_out(n, pad, ksz, stride):
        return -((-(n + 2 * pad - ksz)) // stride) + 1

    def maxpool(self, x, ksz, pad, stride, stream=None):
        import torch
        N, H, W, Cc = x.shape
        y = torch.empty((N, self._pool_out(H, pad, ksz, stride), self._pool_out(W, pad, ksz, stride), Cc),
                        dtype=torch.float32, device=x.device)
        _check(lib.qcnn_maxpool(self.h, _dptr(x), _dptr(y), N, H, W, Cc, ksz, pad, stride, _stream(stream)))
        return y

    def lrn_maxpool(self, x, size, alpha, beta, k, ksz, pad, stride, stream=None):
        import torch
        N, H, W, Cc = x.shape
        y = torch.empty((N, self._pool_out(H, pad, ksz, stride), self._pool_out(W, pad, ksz, stride), Cc),
                        dtype=torch.float32, device=x.device)
        _check(lib.qcnn_lrn_maxpool(self.h, _dptr(x), _dptr(y), N, H, W, Cc, size, alpha, beta, k, ksz, pad, stride,
                                    _stream(stream)))
        return y

    def softmax(self, x, stream=None):
        import torch
        y = torch.empty_like(x)
        _check(lib.qcnn_softmax(self.h, _dptr(x), _dptr(y), x.shape[0], x.numel() // x.shape[0], _stream(stream)))
        return y

    def topk(self, prob, k, mode=0, stream=None):
        """k-fold arg-max (first maximum wins, winner zeroed) of every row; returns (idx int32 [N,k], val [N,k])."""
        import torch
        N, Cc = prob.shape
        idx = torch.empty((N, k), dtype=torch.int32, device=prob.device)
        val = torch.empty((N, k), dtype=torch.float32, device=prob.device)
        _check(lib.qcnn_topk(self.h, _dptr(prob), N, Cc, k, mode, C.c_void_p(idx.data_ptr()), _dptr(val), _stream(stream)))
        return idx, val

    def nchw_to_nhwc(self, x, stream=None):
        import torch
        N, Cc, H, W = x.shape
        y = torch.empty((N, H, W, Cc), dtype=torch.float32, device=x.device)
        _check(lib.qcnn_nchw_to_nhwc(self.h, _dptr(x), _dptr(y), N, Cc, H, W, _stream(stream)))
        return y

    def nhwc_to_nchw(self, x, stream=None):
        import torch
        N, H, W, Cc = x.shape
        y = torch.empty((N, Cc, H, W), dtype=torch.float32, device=x.device)
        _check(lib.qcnn_nhwc_to_nchw(self.h, _dptr(x), _dptr(y), N, H, W, Cc, _stream(stream)))
        return y


class _Layer(object):
    def __init__(self, ctx, h, owned=True):
        self.ctx, self.h, self.owned = ctx, h, owned

    def close(self):
        if self.h and self.owned:
            lib.qcnn_layer_destroy(self.h)
        self.h = None

    def out_dims(self):
        d = (_i * 3)()
        _check(lib.qcnn_layer_out_dims(self.h, d))
        return tuple(d)

    def work(self, N):
        b, l, m = C.c_double(), C.c_double(), C.c_double()
        _check(lib.qcnn_layer_work(self.h, N, C.byref(b), C.byref(l), C.byref(m)))
        return dict(alg_bytes=b.value, lookups=l.value, lut_macs=m.value)

    def describe(self, N):
        buf = C.create_string_buffer(512)
        _check(lib.qcnn_layer_describe(self.h, N, buf, 512))
        return buf.value.decode()

    def set_param(self, name, value):
        _check(lib.qcnn_layer_set_param(self.h, name.encode(), value))

    def read_asmt(self, n):
        out = np.empty(n, np.uint8)
        _check(lib.qcnn_layer_read_asmt_h(self.h, out.ctypes.data_as(_vp), out.size))
        return out


class ConvLayer(_Layer):
    """PQ convolution; parameters as CaffePara holds them (ctrd [S,K,d], asmt [Cout,k,k,S] 0-based, bias [Cout])."""

    def __init__(self, ctx, Cin, Hin, Win, Cout, ksz, pad, stride, grp, ctrd, asmt, bias):
        ctrd, asmt, bias = _np(ctrd, np.float32), _np(asmt, np.uint8), _np(bias, np.float32)
        S, K, d = ctrd.shape
        assert asmt.shape == (Cout, ksz, ksz, S), asmt.shape
        h = _vp()
        _check(lib.qcnn_conv_layer_create(ctx.h, Cin, Hin, Win, Cout, ksz, pad, stride, grp, S, K, d,
                                          ctrd.ctypes.data_as(_vp), asmt.ctypes.data_as(_vp),
                                          bias.ctypes.data_as(_vp), C.byref(h)))
        _Layer.__init__(self, ctx, h)
        self.shape_in = (Hin, Win, Cin)
        self.n_asmt = asmt.size

    def set_src_nchw(self, enable=True):
        _check(lib.qcnn_conv_layer_set_src_nchw(self.h, int(enable)))

    def forward(self, x, relu=False, stream=None):
        import torch
        ho, wo, co = self.out_dims()
        y = torch.empty((x.shape[0], ho, wo, co), dtype=torch.float32, device=x.device)
        _check(lib.qcnn_conv_aprx_forward(self.h, _dptr(x), x.shape[0], _dptr(y), int(relu), _stream(stream)))
        return y


class FcLayer(_Layer):
    """PQ fully-connected layer; ctrd [S,K,d], asmt [Dout,S] 0-based, bias [Dout]."""

    def __init__(self, ctx, Din, ctrd, asmt, bias):
        ctrd, asmt, bias = _np(ctrd, np.float32), _np(asmt, np.uint8), _np(bias, np.float32)
        S, K, d = ctrd.shape
        Dout = asmt.shape[0]
        assert asmt.shape == (Dout, S)
        h = _vp()
        _check(lib.qcnn_fc_layer_create(ctx.h, Din, Dout, S, K, d, ctrd.ctypes.data_as(_vp),
                                        asmt.ctypes.data_as(_vp), bias.ctypes.data_as(_vp), C.byref(h)))
        _Layer.__init__(self, ctx, h)
        self.Din, self.Dout = Din, Dout
        self.n_asmt = asmt.size

    def set_src_nhwc(self, H, W, Cc):
        _check(lib.qcnn_fc_layer_set_src_nhwc(self.h, H, W, Cc))

    def forward(self, x, relu=False, stream=None):
        import torch
        N = x.shape[0]
        y = torch.empty((N, self.Dout), dtype=torch.float32, device=x.device)
        _check(lib.qcnn_fc_aprx_forward(self.h, _dptr(x), N, _dptr(y), int(relu), _stream(stream)))
        return y


def fc_chain_forward(layers, relu, x, stamps=None, stream=None):
    """Consecutive FC layers (FcLayer objects) as one persistent launch; x [N<=4, Din]; returns [N, Dout of the last]."""
    import torch
    n = len(layers)
    hs = (_vp * n)(*[l.h for l in layers])
    rl = (_i * n)(*[int(r) for r in relu])
    N = x.shape[0]
    y = torch.empty((N, layers[-1].out_dims()[2]), dtype=torch.float32, device=x.device)
    sp = C.c_void_p(stamps.data_ptr()) if stamps is not None else None
    _check(lib.qcnn_fc_chain_forward(hs, rl, n, _dptr(x), N, _dptr(y), sp, _stream(stream)))
    return y


class Net(object):
    """Device-resident network == CaffeEva::LoadCaffePara + ExecForwardPass (reference src/CaffeEva.cc:109-261)."""

    def __init__(self, ctx, dirpath, pfx, model="AlexNet", layers=None, in_chw=None):
        h = _vp()
        if layers is None:
            _check(lib.qcnn_net_create(ctx.h, model.encode(), dirpath.encode(), pfx.encode(), C.byref(h)))
        else:
            arr = (LayerInfo * len(layers))(*layers)
            _check(lib.qcnn_net_create_custom(ctx.h, len(layers), arr, in_chw[0], in_chw[1], in_chw[2],
                                              dirpath.encode(), pfx.encode(), C.byref(h)))
        self.h, self.ctx = h, ctx
        self.out_len = lib.qcnn_net_out_len(h)
        self.layer_count = lib.qcnn_net_layer_count(h)

    def close(self):
        if self.h:
            lib.qcnn_net_destroy(self.h)
            self.h = None

    def set_keep_maps(self, keep):
        _check(lib.qcnn_net_set_keep_maps(self.h, int(keep)))

    def set_profiling(self, on):
        _check(lib.qcnn_net_set_profiling(self.h, int(on)))

    def set_chunk(self, n):
        _check(lib.qcnn_net_set_chunk(self.h, n))

    def forward(self, img, prob=None, logits=None, stream=None):
        """img: CUDA float32 [N,C,H,W]; returns prob [N,out_len] (device)."""
        import torch
        N = img.shape[0]
        if prob is None:
            prob = torch.empty((N, self.out_len), dtype=torch.float32, device=img.device)
        _check(lib.qcnn_net_forward(self.h, _dptr(img), N, _dptr(prob), _dptr(logits) if logits is not None else None,
                                    _stream(stream)))
        return prob

    def forward_host(self, img_h, prob_h=None, logits_h=None):
        """img_h: host float32 array/tensor [N,C,H,W] (pinned for full overlap); returns host probs."""
        def ptr(a):
            if hasattr(a, "data_ptr"):
                return C.c_void_p(a.data_ptr())
            return a.ctypes.data_as(_vp)
        N = img_h.shape[0]
        if prob_h is None:
            prob_h = np.empty((N, self.out_len), np.float32)
        _check(lib.qcnn_net_forward_h(self.h, ptr(img_h), N, ptr(prob_h), ptr(logits_h) if logits_h is not None else None))
        return prob_h

    def set_input_mean(self, mean_chw):
        if mean_chw is None:
            _check(lib.qcnn_net_set_input_mean(self.h, None))
        else:
            m = _np(mean_chw, np.float32)
            _check(lib.qcnn_net_set_input_mean(self.h, m.ctypes.data_as(_vp)))

    def forward_u8(self, img_u8, prob=None, logits=None, stream=None):
        """img_u8: CUDA uint8 [N,H,W,C] interleaved pixels; returns prob [N,out_len] (device)."""
        import torch
        N = img_u8.shape[0]
        if prob is None:
            prob = torch.empty((N, self.out_len), dtype=torch.float32, device=img_u8.device)
        assert img_u8.dtype == torch.uint8 and img_u8.is_contiguous()
        _check(lib.qcnn_net_forward_u8(self.h, C.c_void_p(img_u8.data_ptr()), N, _dptr(prob),
                                       _dptr(logits) if logits is not None else None, _stream(stream)))
        return prob

    @staticmethod
    def _hptr(a):
        if a is None:
            return None
        return C.c_void_p(a.data_ptr()) if hasattr(a, "data_ptr") else a.ctypes.data_as(_vp)

    def forward_topk_host(self, img_h, k, mode=0, idx_h=None, val_h=None, prob_h=None):
        """fp32 host images in, on-device top-k out: (idx int32 [N,k], prob [N,k])."""
        N = img_h.shape[0]
        idx_h = np.empty((N, k), np.int32) if idx_h is None else idx_h
        val_h = np.empty((N, k), np.float32) if val_h is None else val_h
        _check(lib.qcnn_net_forward_topk_h(self.h, self._hptr(img_h), N, k, mode, self._hptr(idx_h), self._hptr(val_h),
                                           self._hptr(prob_h)))
        return idx_h, val_h

    def forward_u8_host(self, img_h, k=0, mode=0, idx_h=None, val_h=None, prob_h=None):
        """uint8 host pixels [N,H,W,C] in; top-k (k > 0) and / or probabilities (prob_h) out."""
        N = img_h.shape[0]
        if k > 0:
            idx_h = np.empty((N, k), np.int32) if idx_h is None else idx_h
            val_h = np.empty((N, k), np.float32) if val_h is None else val_h
        elif prob_h is None:
            prob_h = np.empty((N, self.out_len), np.float32)
        _check(lib.qcnn_net_forward_u8_h(self.h, self._hptr(img_h), N, k, mode, self._hptr(idx_h), self._hptr(val_h),
                                         self._hptr(prob_h)))
        return (idx_h, val_h) if k > 0 else prob_h

    def submit_u8_host(self, img_h, k=0, mode=0, idx_h=None, val_h=None, prob_h=None):
        """Asynchronous uint8 step (pinned buffers); returns a ticket for wait()."""
        t = _i()
        _check(lib.qcnn_net_submit_u8_h(self.h, self._hptr(img_h), img_h.shape[0], k, mode, self._hptr(idx_h), self._hptr(val_h),
                                        self._hptr(prob_h), C.byref(t)))
        return t.value

    def wait(self, ticket):
        _check(lib.qcnn_net_wait(self.h, ticket))

    def featmap(self, idx, N):
        """featMapLst[idx] of the last forward as a torch view [N,H,W,C] (None if fused away)."""
        import torch
        p = _vp()
        d = (_i * 4)()
        _check(lib.qcnn_net_featmap(self.h, idx, C.byref(p), d))
        if not p.value:
            return None
        class _View(object):
            pass
        v = _View()
        v.__cuda_array_interface__ = dict(shape=(N, d[1], d[2], d[3]), typestr="<f4", data=(p.value, False), version=2)
        return torch.as_tensor(v, device="cuda:%d" % self.ctx.device).clone()

    def layer_time_ms(self, l):
        v = _f()
        _check(lib.qcnn_net_layer_time_ms(self.h, l, C.byref(v)))
        return v.value

    def layer_work(self, l, N):
        b, lk, m = C.c_double(), C.c_double(), C.c_double()
        _check(lib.qcnn_net_layer_work(self.h, l, N, C.byref(b), C.byref(lk), C.byref(m)))
        return dict(alg_bytes=b.value, lookups=lk.value, lut_macs=m.value)

    def launch_count(self):
        return lib.qcnn_net_launch_count(self.h)

    def pq_layer(self, l):
        h = lib.qcnn_net_pq_layer(self.h, l)
        return _Layer(self.ctx, _vp(h), owned=False) if h else None


class Preproc(object):
    """BmpImgIO on the device (ReszImg, RmMeanImg, CropImg of decoded BMP pixels), qcnn_preproc_*."""

    def __init__(self, ctx, mean_chw, hei_full=256, wid_full=256, hei_crop=227, wid_crop=227, resz_type=0, mean_type=0):
        m = _np(mean_chw, np.float32)
        assert m.ndim == 3 and m.shape[0] == 3
        h = _vp()
        _check(lib.qcnn_preproc_create(ctx.h, resz_type, mean_type, hei_full, wid_full, hei_crop, wid_crop,
                                       m.ctypes.data_as(_vp), m.shape[1], m.shape[2], C.byref(h)))
        self.h, self.ctx, self.crop = h, ctx, (hei_crop, wid_crop)

    def close(self):
        if self.h:
            lib.qcnn_preproc_destroy(self.h)
            self.h = None

    def run(self, images, stream=None):
        """images: list of uint8 numpy arrays [H,W,3] (B,G,R interleaved, top row first); returns CUDA f32 [N,3,hc,wc]."""
        import torch
        N = len(images)
        offs, total = [], 0
        for im in images:
            offs.append(total)
            total += im.size
        flat = np.concatenate([np.ascontiguousarray(im, np.uint8).reshape(-1) for im in images])
        pix = torch.from_numpy(flat).to("cuda:%d" % self.ctx.device)
        out = torch.empty((N, 3) + self.crop, dtype=torch.float32, device=pix.device)
        off = (C.c_longlong * N)(*offs)
        hei = (_i * N)(*[im.shape[0] for im in images])
        wid = (_i * N)(*[im.shape[1] for im in images])
        _check(lib.qcnn_preproc_run(self.h, C.c_void_p(pix.data_ptr()), off, hei, wid, N, _dptr(out), _stream(stream)))
        return out


class MultiNet(object):
    """One process, several GPUs: batch-sharded replicas + NCCL all-gather of the probabilities (qcnn_multi_*)."""

    def __init__(self, dirpath, pfx, model="AlexNet", devices=(0,)):
        devices = list(devices)
        arr = (_i * len(devices))(*devices)
        h = _vp()
        _check(lib.qcnn_multi_create(len(devices), arr, model.encode(), dirpath.encode(), pfx.encode(), C.byref(h)))
        self.h, self.devices = h, devices
        self.out_len = lib.qcnn_multi_out_len(h)

    def close(self):
        if self.h:
            lib.qcnn_multi_destroy(self.h)
            self.h = None

    @property
    def nccl_version(self):
        return lib.qcnn_multi_nccl_version(self.h)

    def pq_layer(self, rank, l):
        net = lib.qcnn_multi_net(self.h, rank)
        h = lib.qcnn_net_pq_layer(net, l)
        return _Layer(None, _vp(h), owned=False) if h else None

    def forward_host(self, img_h, prob_h=None):
        def ptr(a):
            return C.c_void_p(a.data_ptr()) if hasattr(a, "data_ptr") else a.ctypes.data_as(_vp)
        N = img_h.shape[0]
        if prob_h is None:
            prob_h = np.empty((N, self.out_len), np.float32)
        _check(lib.qcnn_multi_forward_h(self.h, ptr(img_h), N, ptr(prob_h)))
        return prob_h

    def shard(self, N, r):
        per = -(-N // len(self.devices))
        return min(N, r * per), min(N, (r + 1) * per)

    def forward(self, shards, N):
        """shards[r]: CUDA float32 tensor on device r with rank r's images; returns device pointers of the gathered probs."""
        R = len(self.devices)
        ins = (_vp * R)(*[(_dptr(t).value if t is not None else None) for t in shards])
        outs = (_vp * R)()
        _check(lib.qcnn_multi_forward(self.h, ins, N, outs))
        return [outs[r] for r in range(R)]

    def sync(self):
        _check(lib.qcnn_multi_sync(self.h))

    def gathered(self, ptr, N, rank):
        """torch view [N, out_len] of a gathered buffer returned by forward() (after sync)."""
        import torch

        class _View(object):
            pass
        v = _View()
        v.__cuda_array_interface__ = dict(shape=(N, self.out_len), typestr="<f4", data=(int(ptr), False), version=2)
        return torch.as_tensor(v, device="cuda:%d" % self.devices[rank])


# ---- file formats (host) ----------------------------------------------------------------------------------
def read_bin_f32(path):
    dc, dims = _i(), (_i * 4)()
    n = lib.qcnn_read_bin_f32(path.encode(), C.byref(dc), dims, None, 0)
    if n < 0:
        raise QcnnError(last_error())
    out = np.empty(n, np.float32)
    lib.qcnn_read_bin_f32(path.encode(), C.byref(dc), dims, out.ctypes.data_as(_vp), n)
    return out.reshape([dims[k] for k in range(dc.value)])


def write_bin_f32(path, arr):
    arr = _np(arr, np.float32)
    dims = (_i * arr.ndim)(*arr.shape)
    _check(lib.qcnn_write_bin_f32(path.encode(), arr.ndim, dims, arr.ctypes.data_as(_vp)))


def read_cbn_u8(path):
    dc, dims, bits = _i(), (_i * 4)(), _i()
    n = lib.qcnn_read_cbn_u8(path.encode(), C.byref(dc), dims, C.byref(bits), None, 0)
    if n < 0:
        raise QcnnError(last_error())
    out = np.empty(n, np.uint8)
    lib.qcnn_read_cbn_u8(path.encode(), C.byref(dc), dims, C.byref(bits), out.ctypes.data_as(_vp), n)
    return out.reshape([dims[k] for k in range(dc.value)]), bits.value


def write_cbn_u8(path, idx0, bits):
    idx0 = _np(idx0, np.uint8)
    dims = (_i * idx0.ndim)(*idx0.shape)
    _check(lib.qcnn_write_cbn_u8(path.encode(), idx0.ndim, dims, idx0.ctypes.data_as(_vp), bits))
