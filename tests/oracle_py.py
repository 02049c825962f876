"""ctypes face of oracle/libgrab_oracle.so (TEST INFRASTRUCTURE; see oracle/grab_oracle.h)."""
import ctypes
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if not os.path.exists(os.path.join(ROOT, "oracle", "libgrab_oracle.so")):  # test infrastructure: gcc, <1 s
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=True, stdout=subprocess.DEVNULL)
_lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "libgrab_oracle.so"))
_libc = ctypes.CDLL(None)

MODE_ALL, MODE_FIRST, MODE_LINE = 0, 1, 2
GO_LITERAL = 1


class _Match(ctypes.Structure):
    _fields_ = [("start", ctypes.c_uint64), ("len", ctypes.c_uint32), ("unit", ctypes.c_uint32)]


class _Matches(ctypes.Structure):
    _fields_ = [("v", ctypes.POINTER(_Match)), ("n", ctypes.c_size_t), ("cap", ctypes.c_size_t)]


class _Opts(ctypes.Structure):
    _fields_ = [("print_offset", ctypes.c_int), ("print_line", ctypes.c_int), ("single", ctypes.c_int),
                ("colored", ctypes.c_int), ("strict_q2", ctypes.c_int), ("path_prefix", ctypes.c_char_p),
                ("chunk_size", ctypes.c_size_t)]


_lib.go_compile.restype = ctypes.c_void_p
_lib.go_compile.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint, ctypes.c_char_p, ctypes.c_size_t]
_lib.go_free.argtypes = [ctypes.c_void_p]
for _f in ("go_minlen", "go_capture_count", "go_nullable"):
    getattr(_lib, _f).argtypes = [ctypes.c_void_p]
    getattr(_lib, _f).restype = ctypes.c_int
_lib.go_scan_window.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint32,
                                ctypes.c_int, ctypes.c_int, ctypes.POINTER(_Matches)]
_lib.go_scan_window.restype = ctypes.c_int
_lib.go_matches_free.argtypes = [ctypes.POINTER(_Matches)]
_lib.go_grab_buffer.argtypes = [ctypes.c_void_p, ctypes.POINTER(_Opts), ctypes.c_char_p, ctypes.c_size_t,
                                ctypes.c_void_p]
_lib.go_grab_buffer.restype = ctypes.c_int
_libc.open_memstream.restype = ctypes.c_void_p
_libc.open_memstream.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]
_libc.fclose.argtypes = [ctypes.c_void_p]
_libc.free.argtypes = [ctypes.c_void_p]


class OracleError(Exception):
    pass


class Regex:
    def __init__(self, pattern, literal=False):
        if isinstance(pattern, str):
            pattern = pattern.encode("latin-1")
        err = ctypes.create_string_buffer(256)
        self._h = _lib.go_compile(pattern, len(pattern), GO_LITERAL if literal else 0, err, 256)
        if not self._h:
            raise OracleError(err.value.decode())

    def __del__(self):
        if getattr(self, "_h", None):
            _lib.go_free(self._h)
            self._h = None

    @property
    def minlen(self):
        return _lib.go_minlen(self._h)

    @property
    def captures(self):
        return _lib.go_capture_count(self._h)

    @property
    def nullable(self):
        return bool(_lib.go_nullable(self._h))

    def scan_window(self, data, base_off=0, unit=0, mode=MODE_ALL, strict_q2=True):
        """[(absolute start, len)] for one scan unit (grab.cc:175-213)."""
        data = bytes(data)
        m = _Matches()
        rc = _lib.go_scan_window(self._h, data, len(data), base_off, unit, mode, int(strict_q2), ctypes.byref(m))
        out = [(m.v[i].start, m.v[i].len) for i in range(m.n)]
        _lib.go_matches_free(ctypes.byref(m))
        if rc < 0:
            raise OracleError("scan error (empty-matchable pattern?)")
        return out

    def scan_window_np(self, data, base_off=0, mode=MODE_ALL, strict_q2=True):
        """scan_window() for big windows: `data` a contiguous numpy uint8 array (not copied), the result a structured numpy
        array with fields start / len / unit (millions of matches without millions of Python tuples)."""
        import numpy as np
        a = np.ascontiguousarray(data, dtype=np.uint8)
        m = _Matches()
        rc = _lib.go_scan_window(self._h, ctypes.cast(a.ctypes.data, ctypes.c_char_p), a.size, base_off, 0, mode, int(strict_q2), ctypes.byref(m))
        dt = np.dtype([("start", "<u8"), ("len", "<u4"), ("unit", "<u4")])
        out = np.frombuffer(ctypes.string_at(m.v, m.n * dt.itemsize), dtype=dt).copy() if m.n else np.zeros(0, dtype=dt)
        _lib.go_matches_free(ctypes.byref(m))
        if rc < 0:
            raise OracleError("scan error (empty-matchable pattern?)")
        return out

    def scan_file(self, data, chunk_size=1 << 30, mode=MODE_ALL, strict_q2=True):
        """All windows of grab.cc:154-159 over an in-memory file; duplicates in overlaps are kept (Q3).
        FIRST mode stops at the first window that printed something (grab.cc:232-233)."""
        data = bytes(data)
        if self.minlen > len(data):
            return []
        out, off = [], 0
        while off < len(data):
            clen = min(chunk_size, len(data) - off)
            r = self.scan_window(data[off:off + clen], off, 0, mode, strict_q2)
            out += r
            if r and mode == MODE_FIRST:
                break
            off += chunk_size - 4096
        return out

    def grab(self, data, offsets=False, line=True, single=False, colored=False, strict_q2=True, path=None,
             chunk_size=1 << 30):
        """Exactly the bytes the reference prints for one file."""
        data = bytes(data)
        o = _Opts(int(offsets), int(line), int(single), int(colored), int(strict_q2),
                  path.encode() if path is not None else None, chunk_size)
        bufp, sz = ctypes.c_void_p(), ctypes.c_size_t()
        fp = _libc.open_memstream(ctypes.byref(bufp), ctypes.byref(sz))
        rc = _lib.go_grab_buffer(self._h, ctypes.byref(o), data, len(data), fp)
        _libc.fclose(fp)
        out = ctypes.string_at(bufp.value, sz.value)
        _libc.free(bufp)
        if rc < 0:
            raise OracleError("grab error")
        return out
