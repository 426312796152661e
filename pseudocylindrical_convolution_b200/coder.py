"""`coder` - mirror of the reference's arithmetic-coder module (coder/python.cpp:63-72) over libpcx's host
range coder (csrc/pcx_coder.cpp).  Same class name and the same eight methods; the bitstream is
byte-identical to the reference's (tests/test_oracle_cpu.py compares against the compiled reference coder)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import PcxError, call


def _int32_cpu(t, name):
    if isinstance(t, torch.Tensor):
        if t.is_cuda or t.dtype != torch.int32:
            t = t.to("cpu").to(torch.int32)          # my_encoder() converts; encodes() expects int32 CPU (python.cpp:22-24)
        if not t.is_contiguous():
            t = t.contiguous()
        return t, C.c_void_p(t.data_ptr())
    a = np.ascontiguousarray(t, dtype=np.int32)
    return a, C.c_void_p(a.ctypes.data)


class coder:
    """coder.coder(path)"""

    def __init__(self, name):
        _lib.load()
        self._name = str(name)
        self._h = _lib.load().pcx_coder_open(self._name.encode())
        if not self._h:
            raise PcxError("cannot create coder")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().pcx_coder_close(h)
            except Exception:
                pass

    # --- encoder
    def start_encoder(self):
        call("pcx_coder_start_encoder", self._h)

    def encodes(self, table, ncode, symbols, num):
        """encodes(table int32 (n, ncode+1), ncode, symbols int32 (n,), num) - python.cpp:22-41"""
        tk, tp = _int32_cpu(table, "table")
        sk, sp = _int32_cpu(symbols, "symbols")
        call("pcx_coder_encodes", self._h, tp, int(ncode), sp, int(num))

    def encode(self, table, ncode, total, symbol):
        """scalar form - python.cpp:4-12"""
        tk, tp = _int32_cpu(table, "table")
        sym = np.array([int(symbol)], np.int32)
        call("pcx_coder_encodes", self._h, tp, int(ncode), C.c_void_p(sym.ctypes.data), 1)

    def end_encoder(self):
        call("pcx_coder_end_encoder", self._h)

    # --- decoder
    def start_decoder(self):
        call("pcx_coder_start_decoder", self._h)

    def decodes(self, table, ncode, num):
        """decodes(table, ncode, num) -> float32 CPU tensor with table.size(0) entries, first num valid (python.cpp:42-60)"""
        tk, tp = _int32_cpu(table, "table")
        rows = int(tk.shape[0])
        out = torch.empty(rows, dtype=torch.float32)
        call("pcx_coder_decodes", self._h, tp, int(ncode), int(num), C.c_void_p(out.data_ptr()))
        return out

    def decode(self, table, ncode, total):
        tk, tp = _int32_cpu(table, "table")
        out = np.zeros(1, np.float32)
        call("pcx_coder_decodes", self._h, tp, int(ncode), 1, C.c_void_p(out.ctypes.data))
        return int(out[0])

    # --- in-memory variants (B200 pipeline: bitstreams stay in host RAM)
    def start_encoder_mem(self):
        call("pcx_coder_start_encoder_mem", self._h)

    def take_bytes(self):
        n = call("pcx_coder_take_bytes", self._h, None, 0)
        buf = (C.c_ubyte * max(n, 1))()
        call("pcx_coder_take_bytes", self._h, buf, n)
        return bytes(buf[:n])

    def start_decoder_mem(self, data):
        data = bytes(data)
        call("pcx_coder_start_decoder_mem", self._h, data, len(data))
