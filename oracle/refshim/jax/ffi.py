"""jax.ffi stand-in (test infrastructure): register_ffi_target / pycapsule / ffi_call with the call shape of the real
API, dispatching to handlers compiled against tests/mock_xla/xla/ffi/api/ffi.h.  A call does what XLA does around a
custom call: places the operands on the device, allocates the result buffers (or aliases a donated operand,
input_output_aliases), passes the compute stream, attributes by name and the buffers in order, and raises if the handler
returns an error.  Arrays are torch tensors, like everywhere in this stand-in."""
import ctypes

import numpy as _np
import torch

_targets = {}


class MockBuffer(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("rank", ctypes.c_int32), ("dims", ctypes.c_int64 * 6)]


class MockAttr(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char_p), ("is_double", ctypes.c_int32), ("d", ctypes.c_double), ("i", ctypes.c_int32)]


class MockCallFrame(ctypes.Structure):
    _fields_ = [("stream", ctypes.c_void_p), ("nargs", ctypes.c_int32), ("args", ctypes.POINTER(MockBuffer)),
                ("nrets", ctypes.c_int32), ("rets", ctypes.POINTER(MockBuffer)), ("nattrs", ctypes.c_int32),
                ("attrs", ctypes.POINTER(MockAttr)), ("error", ctypes.c_char * 256)]


def pycapsule(fn):
    return fn


def register_ffi_target(name, capsule, platform="cpu", **kwargs):
    capsule.restype = ctypes.c_int
    capsule.argtypes = [ctypes.POINTER(MockCallFrame)]
    _targets[name] = (capsule, platform)


def _device():
    return "cuda" if torch.cuda.is_available() else "cpu"


def _stream():
    try:
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    except Exception:
        return None


def _buffers(tensors):
    arr = (MockBuffer * max(len(tensors), 1))()
    for k, t in enumerate(tensors):
        assert t.dim() <= 6 and t.is_contiguous()
        arr[k].data, arr[k].rank = t.data_ptr(), t.dim()
        for d, s in enumerate(t.shape):
            arr[k].dims[d] = s
    return arr


def ffi_call(target_name, result_shape_dtypes, input_output_aliases=None, **kwargs):
    single = not isinstance(result_shape_dtypes, (tuple, list))
    outs = [result_shape_dtypes] if single else list(result_shape_dtypes)
    aliases = dict(input_output_aliases or {})

    def call(*args, **attrs):
        fn, _ = _targets[target_name]
        dev = _device()
        ins = [torch.as_tensor(a).to(dev).contiguous() for a in args]
        res = []
        for k, sd in enumerate(outs):
            donated = [i for i, o in aliases.items() if o == k]
            if donated:
                t = ins[donated[0]]
                assert tuple(t.shape) == tuple(sd.shape) and t.dtype == sd.dtype, "aliased operand / result mismatch"
                res.append(t)
            else:
                res.append(torch.empty(tuple(int(s) for s in sd.shape), dtype=sd.dtype, device=dev))
        names = sorted(attrs)
        at = (MockAttr * max(len(names), 1))()
        keep = []
        for k, n in enumerate(names):
            v = attrs[n]
            b = n.encode()
            keep.append(b)
            at[k].name = b
            if isinstance(v, (float, _np.floating)):
                at[k].is_double, at[k].d = 1, float(v)
            else:
                at[k].is_double, at[k].i = 0, int(v)
        a_in, a_out = _buffers(ins), _buffers(res)
        frame = MockCallFrame(_stream(), len(ins), a_in, len(res), a_out, len(names), at, b"")
        rc = fn(ctypes.byref(frame))
        if rc:
            raise RuntimeError("XLA FFI call %s failed: %s" % (target_name, frame.error.decode(errors="replace")))
        return res[0] if single else tuple(res)
    return call
