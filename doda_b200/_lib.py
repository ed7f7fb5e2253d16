"""ctypes binding of libb200sparse.so (the C ABI declared in include/b200sparse.h).

The prototypes are parsed from the header itself so the Python side cannot drift from the C ABI.
There is NO fallback: if the shared library is missing, importing an op raises; if a call returns a
non-zero code, a RuntimeError carrying b200sp_last_error() is raised.
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(_HERE, "..", "include", "b200sparse.h")
LIB_PATH = os.path.join(_HERE, "libb200sparse.so")

_CTYPE = {
    "int": ctypes.c_int,
    "int32_t": ctypes.c_int32,
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
}


def parse_header(path=HEADER):
    """-> {name: (restype, [(argname, ctype), ...])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"(const\s+char\s*\*|int64_t|int)\s+(b200sp_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = ctypes.c_char_p if "char" in ret else _CTYPE[ret]
        argl = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    ct = ctypes.c_void_p
                    an = a.split("*")[-1].strip()
                else:
                    toks = a.replace("const ", "").split()
                    ct = _CTYPE[toks[0]]
                    an = toks[-1]
                argl.append((an, ct))
        protos[name] = (restype, argl)
    return protos


class _Lib:
    def __init__(self):
        self._dll = None
        self.protos = None

    def load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libb200sparse.so not found at %s -- build it with `python -m doda_b200.build` "
                "(there is no CPU/PyTorch fallback for the sm_100a kernels)" % LIB_PATH)
        dll = ctypes.CDLL(LIB_PATH)
        self.protos = parse_header()
        for name, (restype, argl) in self.protos.items():
            fn = getattr(dll, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype = restype
            fn.argtypes = [ct for _, ct in argl]
        self._dll = dll
        return dll

    def __getattr__(self, name):
        dll = self.load()
        fn = getattr(dll, name)
        setattr(self, name, fn)
        return fn


lib = _Lib()


def last_error():
    return lib.b200sp_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("libb200sparse %s failed (code %d): %s" % (what, rc, last_error()))
