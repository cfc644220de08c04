"""alevin ``bfh.txt`` -> the arrays of the quant path (host side of ``scasa_bfh_*``).

Replaces scasa_quant_step1_1_v1.0.0.R (:16-68) and the ordering at scasa_quant_step2_v1.0.0.R:59-60:
no exploded (cell, eqclass, transcript) table, no ``Sample_eqClass.RData`` round trip.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import lib as _l


@dataclass
class Bfh:
    tx_names: list          # the file's transcript list (step1_1:19)
    barcodes: list          # cells = barcodes that occur, byte order (step2:59-60)
    eq_off: np.ndarray      # int64 [n_eq+1]
    eq_tx: np.ndarray       # int32, indices into tx_names
    cell_ptr: np.ndarray    # int64 [n_cells+1]
    eq_id: np.ndarray       # int32, 0-based eqclass line
    eq_count: np.ndarray    # int32, number of UMIs (step1_1:50-56)

    def tx_ids(self, xm) -> np.ndarray:
        """eq_tx in the X-matrix's transcript dictionary (what ``scasa_eqclass_map`` wants);
        transcripts the X-matrix does not know get ids past its dictionary."""
        lut = getattr(xm, "_tx_lut", None)                 # the X-matrix's dictionary, built once per X-matrix
        if lut is None:
            lut = {}
            for i, name in enumerate(xm.tx_names):
                lut[name] = i                               # a repeated name keeps its last index, as the dict comprehension did
            try:
                xm._tx_lut = lut
            except AttributeError:
                pass
        n = len(xm.tx_names)
        table = np.array([lut.get(name, n + k) for k, name in enumerate(self.tx_names)], np.int32)
        return table[self.eq_tx]


def read_bus(bus_txt: str, matrix_ec: str, transcripts_txt: str, libsize_thres: int = 200) -> Bfh:
    """kallisto | bustools output -> the same arrays (replaces scasa_quant_step1_2_v1.0.0.R):
    ``eq_id`` indexes the lines of matrix.ec, ``eq_count`` is the number of kept molecules."""
    L = _l.load()
    h = C.c_void_p()
    rc = L.scasa_bus_open(bus_txt.encode(), matrix_ec.encode(), transcripts_txt.encode(), libsize_thres, C.byref(h))
    if rc:
        raise _l.ScasaError(rc, L.scasa_last_error(None).decode())
    return _collect(L, h)


def read_bfh(path: str, n_threads: int = 0) -> Bfh:
    L = _l.load()
    h = C.c_void_p()
    rc = L.scasa_bfh_open(path.encode(), n_threads, C.byref(h))
    if rc:
        raise _l.ScasaError(rc, L.scasa_last_error(None).decode())
    return _collect(L, h)


def _collect(L, h) -> Bfh:
    try:
        s = [C.c_int64(0) for _ in range(7)]
        L.scasa_bfh_sizes(h, *[C.byref(x) for x in s])
        n_tx, n_cells, n_eq, n_pairs, n_eq_tx, txb, cbb = (x.value for x in s)
        tx = C.create_string_buffer(max(txb, 1))
        cb = C.create_string_buffer(max(cbb, 1))
        eq_off = np.zeros(n_eq + 1, np.int64)
        eq_tx = np.zeros(max(n_eq_tx, 1), np.int32)
        cell_ptr = np.zeros(n_cells + 1, np.int64)
        eq_id = np.zeros(max(n_pairs, 1), np.int32)
        eq_count = np.zeros(max(n_pairs, 1), np.int32)
        rc = L.scasa_bfh_read(h, tx, cb, _l.ptr(eq_off, C.c_int64), _l.ptr(eq_tx, C.c_int32), _l.ptr(cell_ptr, C.c_int64),
                              _l.ptr(eq_id, C.c_int32), _l.ptr(eq_count, C.c_int32))
        if rc:
            raise _l.ScasaError(rc, L.scasa_last_error(None).decode())
        names = tx.raw[:txb].decode().split("\n")[:-1] if txb else []
        codes = cb.raw[:cbb].decode().split("\n")[:-1] if cbb else []
        return Bfh(names, codes, eq_off, eq_tx[:n_eq_tx], cell_ptr, eq_id[:n_pairs], eq_count[:n_pairs])
    finally:
        L.scasa_bfh_close(h)
