"""ctypes binding of libdsvgp_b200.so (the C ABI declared in include/dsvgp_b200.h).

The prototypes are parsed from the header itself, so the header is the single source of truth: every symbol it
declares must be exported by the library (tests/test_abi.py checks this without a GPU).  There is NO fallback:
if the library is missing the import raises, and every compute call raises unless it returns DSVGP_OK.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdsvgp_b200.so")
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "include", "dsvgp_b200.h"))

_CTYPES = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "double": ctypes.c_double, "size_t": ctypes.c_size_t,
           "dsvgp_stream_t": ctypes.c_void_p, "void": None}
_PROTO = re.compile(r"^(int|void|size_t|int64_t)\s+(dsvgp_\w+)\s*\(([^;]*)\)\s*;", re.M)

ERRORS = {-1: "DSVGP_ERR_ARG (bad argument)", -2: "DSVGP_ERR_LAUNCH (CUDA launch failed)",
          -3: "DSVGP_ERR_WORKSPACE (workspace too small)"}


class DsvgpError(RuntimeError):
    pass


def parse_header(path=HEADER_PATH):
    """-> {name: (restype, [argtypes])} for every prototype in the header."""
    protos = {}
    with open(path) as f:
        text = f.read()
    for ret, name, args in _PROTO.findall(text):
        argtypes = []
        args = args.strip()
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a:
                    argtypes.append(ctypes.c_void_p)
                else:
                    argtypes.append(_CTYPES[a.replace("const ", "").split()[0]])
        protos[name] = (_CTYPES[ret], argtypes)
    return protos


PROTOTYPES = parse_header()

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python gp-derivatives-variational-inference_b200/csrc/build.py` "
        "(or __graft_entry__.build()).  dsvgp_b200 has no CPU or PyTorch fallback.")
_lib = ctypes.CDLL(LIB_PATH)
for _name, (_ret, _args) in PROTOTYPES.items():
    _fn = getattr(_lib, _name)      # AttributeError here = header/library mismatch
    _fn.restype = _ret
    _fn.argtypes = _args


def _arg(a):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        if not a.is_cuda:
            raise DsvgpError("dsvgp_b200 kernels take CUDA tensors only (there is no CPU path)")
        return ctypes.c_void_p(a.data_ptr())
    return a


# torch.cuda.current_stream() costs ~14 us of Python per call (device-index resolution, availability checks) and a step makes ~50
# calls: the raw accessors torch itself uses underneath are two C calls.
_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_GET_DEVICE = getattr(torch._C, "_cuda_getDevice", None)


def _current_device():
    return _GET_DEVICE() if _GET_DEVICE is not None else torch.cuda.current_device()


def stream():
    if _RAW_STREAM is not None and _GET_DEVICE is not None:
        return ctypes.c_void_p(_RAW_STREAM(_GET_DEVICE()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    """Call an int-returning entry point on torch's current stream; raise on a negative status.
    Returns the (non-negative) status: a few entry points use 1 to report an optional extra output."""
    # per-device state (streams, function attributes, work lists) follows the CURRENT device: refuse tensors that live elsewhere
    # instead of launching on the wrong device / stream (one process per GPU is the supported layout).  One pass over the
    # arguments: this runs ~200 times per training step.
    conv, dev = [], None
    for a in args:
        if isinstance(a, torch.Tensor):
            if not a.is_cuda:
                raise DsvgpError("dsvgp_b200 kernels take CUDA tensors only (there is no CPU path)")
            if dev is None:
                dev = _current_device()
            if a.device.index != dev:
                raise DsvgpError(f"{name}: tensor on cuda:{a.device.index} but the current device is cuda:{dev} "
                                 "(use torch.cuda.set_device / torch.cuda.device(...) around the call)")
            conv.append(ctypes.c_void_p(a.data_ptr()))
        else:
            conv.append(a)
    rc = getattr(_lib, name)(*conv, stream())
    if rc < 0:
        raise DsvgpError(f"{name} failed: {ERRORS.get(rc, rc)}")
    return rc


def call_raw(name, *args):
    """Entry points without a stream argument / with a non-status return value."""
    return getattr(_lib, name)(*[_arg(a) for a in args])


def suffix(dtype):
    if dtype == torch.float32:
        return "f32"
    if dtype == torch.float64:
        return "f64"
    raise DsvgpError(f"unsupported dtype {dtype}: the hot path is built for float32 and float64")


def chol_plan(Mq):
    Mp, nb0, nlev = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _lib.dsvgp_chol_plan(int(Mq), ctypes.byref(Mp), ctypes.byref(nb0), ctypes.byref(nlev))
    return Mp.value, nb0.value, nlev.value


def version():
    return _lib.dsvgp_version()


def launch_count():
    """CUDA kernels launched by the library so far in this process."""
    return int(_lib.dsvgp_launch_count())
