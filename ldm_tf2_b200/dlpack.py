"""DLPack consumer for the Python shim: unwraps a capsule (tf.experimental.dlpack.to_dlpack,
torch, cupy, numpy ...) into (pointer, shape, dtype, device) so that only plain pointers and
sizes cross the C ABI.  Ownership follows the DLPack contract: the capsule is renamed to
"used_dltensor" and its deleter is called exactly once, after the C call has returned."""
from __future__ import annotations

import ctypes as C

import numpy as np

kDLCPU, kDLCUDA, kDLCUDAHost, kDLCUDAManaged = 1, 2, 3, 13
_CODES = {(0, 32): np.int32, (0, 64): np.int64, (2, 32): np.float32, (1, 8): np.uint8}


class DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", DLDevice), ("ndim", C.c_int32), ("dtype", DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class DLManagedTensor(C.Structure):
    pass


_DELETER = C.CFUNCTYPE(None, C.POINTER(DLManagedTensor))
DLManagedTensor._fields_ = [("dl_tensor", DLTensor), ("manager_ctx", C.c_void_p), ("deleter", _DELETER)]

_api = C.pythonapi
_api.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_IsValid.restype = C.c_int
_api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_GetPointer.restype = C.c_void_p
_api.PyCapsule_SetName.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_SetName.restype = C.c_int


class Borrowed:
    """A tensor borrowed through DLPack.  `ptr` is a host or device address usable by the C ABI
    (which detects the memory space itself).  Call release() after the C call."""

    def __init__(self, capsule):
        if not _api.PyCapsule_IsValid(capsule, b"dltensor"):
            raise ValueError("expected an unconsumed DLPack capsule named 'dltensor'")
        self._managed = C.cast(_api.PyCapsule_GetPointer(capsule, b"dltensor"), C.POINTER(DLManagedTensor))
        _api.PyCapsule_SetName(capsule, b"used_dltensor")
        self._capsule = capsule
        t = self._managed.contents.dl_tensor
        self.device_type, self.device_id = t.device.device_type, t.device.device_id
        if self.device_type not in (kDLCPU, kDLCUDA, kDLCUDAHost, kDLCUDAManaged):
            self.release()
            raise ValueError(f"unsupported DLPack device type {self.device_type}")
        key = (t.dtype.code, t.dtype.bits)
        if t.dtype.lanes != 1 or key not in _CODES:
            self.release()
            raise ValueError(f"unsupported DLPack dtype code={t.dtype.code} bits={t.dtype.bits}")
        self.dtype = np.dtype(_CODES[key])
        self.shape = tuple(t.shape[i] for i in range(t.ndim))
        if t.strides:  # must be compact row-major
            expect = 1
            for i in range(t.ndim - 1, -1, -1):
                if self.shape[i] != 1 and t.strides[i] != expect:
                    self.release()
                    raise ValueError("DLPack tensor must be compact row-major")
                expect *= self.shape[i]
        self.ptr = (t.data or 0) + t.byte_offset
        self.on_device = self.device_type in (kDLCUDA, kDLCUDAManaged)

    def release(self):
        if self._managed is not None:
            m = self._managed
            self._managed = None
            if m.contents.deleter:
                m.contents.deleter(m)

    def __del__(self):
        self.release()


def _device_synchronize():
    """cudaDeviceSynchronize through the library (it links the CUDA runtime statically): a raw
    DLPack capsule carries no stream, so device producers are fenced on the host."""
    from . import lib as _lib
    _lib.check(_lib.load().ldm_device_synchronize())


def borrow(obj, dtype=None):
    """numpy array / DLPack capsule / object with __dlpack__ -> (pointer-or-array, shape, keepalive).
    Host tensors come back as C-contiguous numpy arrays of `dtype`; device tensors as a Borrowed
    whose .ptr is passed straight through the C ABI."""
    if isinstance(obj, np.ndarray) or np.isscalar(obj) or isinstance(obj, (list, tuple)):
        a = np.ascontiguousarray(obj, dtype=dtype)
        return a, a.shape, a
    cap = obj
    if hasattr(obj, "__dlpack__"):
        dev = obj.__dlpack_device__() if hasattr(obj, "__dlpack_device__") else (kDLCPU, 0)
        if int(dev[0]) == kDLCPU:
            a = np.ascontiguousarray(np.from_dlpack(obj), dtype=dtype)
            return a, a.shape, a
        # stream=-1: "producer, do not synchronise for me" is NOT what we want -- the library copies
        # on its own non-blocking stream, so the producer's pending work must be complete.  Ask the
        # producer to make the data safe for the legacy default stream (stream=1), then block the
        # host on the device so that our stream cannot overtake it.
        try:
            cap = obj.__dlpack__(stream=1)
        except TypeError:
            cap = obj.__dlpack__()
        _device_synchronize()
    b = Borrowed(cap)
    if b.on_device and cap is obj:
        _device_synchronize()
    if dtype is not None and b.dtype != np.dtype(dtype):
        b.release()
        raise ValueError(f"expected dtype {np.dtype(dtype)}, got {b.dtype}")
    if not b.on_device:
        n = int(np.prod(b.shape)) if b.shape else 1
        buf = (C.c_char * (n * b.dtype.itemsize)).from_address(b.ptr)
        a = np.frombuffer(buf, dtype=b.dtype).reshape(b.shape).copy()
        b.release()
        return a, a.shape, a
    return b.ptr, b.shape, b
