"""TEST INFRASTRUCTURE ONLY -- golden vectors for the UART TLV decode from the reference's OWN
decoder, `ReadIWR14xx.read()` (ReadDataIWR1443.py:27-201), executed unmodified.

Two things stand between that code and this image:
  * it imports `serial` (pyserial) and talks to two ports -> a stub module whose Serial object hands `read()` the
    packet bytes; the instance is created without running __init__ (no ports, no config upload) and its
    configParameters come from the reference's own `__parseConfigFile` on the reference's own config_cases/*.cfg;
  * it relies on numpy 1.26 casting (requirements.txt:8): out-of-range values stored into int16 arrays wrap like a C
    cast, and `int16_array - 65535` is computed in int32 (value-based promotion) and wraps on assignment.  numpy >= 2
    raises OverflowError for both (NEP 50) -- for EVERY frame, because the subtraction is evaluated even when its mask
    is empty.  The module therefore sees a proxy `np` whose int16 `zeros` return an ndarray subclass that restores
    exactly those two numpy-1.x behaviours (`_Int16Legacy` below) and nothing else.
So header offsets, field order, the Q format, the Doppler correction and the float arithmetic are the reference's; the
integer casting semantics are numpy 1.26's as documented, supplied by the shim.  Output: tests/golden/tlv/reference_tlv.npz
(packets, decoded detObj fields, dataOK, frame numbers).  Run here (needs /root/reference): python oracle/gen_tlv_golden.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import tlv_oracle as to  # noqa: E402

REF = os.environ.get("MMW_REFERENCE_ROOT", "/root/reference")


class _Int16Legacy(np.ndarray):
    """int16 array with numpy 1.x semantics for the two operations ReadIWR14xx.read relies on."""

    def __setitem__(self, key, value):
        v = np.asarray(value)
        if v.dtype.kind in "iu":
            v = v.astype(np.int64).astype(np.int16)              # C cast: wraps modulo 2^16
        super().__setitem__(key, v)

    def __sub__(self, other):
        if isinstance(other, int) and not -32768 <= other <= 32767:
            return np.asarray(self).astype(np.int32) - other     # value-based promotion: int16 (-) 65535 -> int32
        return np.asarray(self).__sub__(other)


class _LegacyNumpy:
    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        return getattr(self._real, name)

    def zeros(self, shape, dtype=float, **kw):
        a = self._real.zeros(shape, dtype=dtype, **kw)
        return a.view(_Int16Legacy) if a.dtype == np.int16 else a


class _Serial:
    def __init__(self, *a, **k):
        self.buf = b""

    @property
    def in_waiting(self):
        return len(self.buf)

    def read(self, n):
        b, self.buf = self.buf[:n], self.buf[n:]
        return b

    def write(self, b):
        pass


def load_reader_module():
    serial = types.ModuleType("serial")
    serial.Serial = _Serial
    sys.modules["serial"] = serial
    sys.path.insert(0, os.path.join(REF, "src"))
    import ReadDataIWR1443 as R
    R.np = _LegacyNumpy(np)
    R.ReadIWR14xx.__del__ = lambda self: None                    # the real one closes serial ports
    return R


def new_reader(R, cfg_file):
    r = object.__new__(R.ReadIWR14xx)
    r.configFileName = cfg_file
    r.byteBuffer = np.zeros(2 ** 15, dtype="uint8")
    r.byteBufferLength = 0
    r.configParameters = r._ReadIWR14xx__parseConfigFile()
    r.MMWDEMO_UART_MSG_DETECTED_POINTS = 1
    r.maxBufferSize = 2 ** 15
    r.magicWord = [2, 1, 4, 3, 6, 5, 8, 7]
    r.Dataport = _Serial()
    return r


def raw_packet(frame, words, q, tlv_type=1, n_det=None, tail=b""):
    """A framed packet from explicit 16-bit words [n, 6] = rangeIdx, dopplerIdx, peakVal, x, y, z (unsigned)."""
    import struct
    words = np.asarray(words, np.int64).reshape(-1, 6)
    n = words.shape[0]
    body = (words & 0xffff).astype("<u2").tobytes()
    tlv = struct.pack("<2I2H", tlv_type, 4 + len(body), n, q) + body
    total = to.HEADER_BYTES + len(tlv) + len(tail)
    hdr = to.MAGIC + struct.pack("<7I", 0x01020003, total, 0xA1443, frame, 0, n if n_det is None else n_det, 1)
    return hdr + tlv + tail


def main():
    R = load_reader_module()
    cfgs = ["iwr1443sdk2_4m_12hz.cfg", "our_config_8.5m.cfg", "mars_config.cfg"]
    rng = np.random.default_rng(1443)
    packets, cfg_idx = [], []
    for ci, cfg in enumerate(cfgs):
        for k in range(6):
            n = int(rng.integers(1, 40))
            words = np.stack([rng.integers(0, 200, n),                            # rangeIdx
                              rng.integers(-20, 20, n),                           # dopplerIdx, two's complement on the wire
                              rng.integers(0, 2000, n),                           # peakVal
                              rng.integers(-2000, 2000, n), rng.integers(0, 4000, n), rng.integers(-800, 800, n)], 1)
            if k == 3:
                words[:, 1] = rng.integers(16, 300, n)        # positive indices above numDopplerBins/2 - 1: the "- 65535" quirk
            if k == 4:
                words[:, 2] = rng.integers(30000, 65536, n)   # peakVal beyond int16: wraps
            packets.append(raw_packet(100 * ci + k + 1, words, 9 if k % 2 == 0 else 7, tail=b"\x00" * (4 * (k % 3))))
            cfg_idx.append(ci)
        packets.append(raw_packet(100 * ci + 50, np.zeros((3, 6)), 9, tlv_type=2)); cfg_idx.append(ci)    # not detected points
        packets.append(raw_packet(100 * ci + 51, np.zeros((2, 6)), 9, n_det=0)); cfg_idx.append(ci)       # numDetectedObj = 0
        packets.append(raw_packet(100 * ci + 52, rng.integers(0, 100, (5, 6)), 9)[:-7]); cfg_idx.append(ci)   # incomplete
    ok_l, fn_l, rows_l, par = [], [], [], []
    for pkt, ci in zip(packets, cfg_idx):
        r = new_reader(R, os.path.join(REF, "config_cases", cfgs[ci]))
        r.Dataport.buf = pkt
        ok, fn, det = r.read()
        ok_l.append(int(ok)); fn_l.append(int(fn))
        par.append((float(r.configParameters["numDopplerBins"]), float(r.configParameters["dopplerResolutionMps"])))
        if ok:
            rows_l.append(np.stack([np.asarray(det[k], np.float64) for k in ("x", "y", "z", "doppler", "peakVal")], 1))
            assert det["x"].dtype == np.float64 and det["doppler"].dtype == np.float64
        else:
            rows_l.append(np.zeros((0, 5)))
    out = os.path.join(ROOT, "tests", "golden", "tlv")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(
        os.path.join(out, "reference_tlv.npz"),
        packet_bytes=np.frombuffer(b"".join(packets), np.uint8),
        packet_offsets=np.cumsum([0] + [len(p) for p in packets]).astype(np.int64),
        cfg=np.array(cfg_idx, np.int32), cfg_names=np.array(cfgs),
        params=np.array(par, np.float64), ok=np.array(ok_l, np.int32), frame=np.array(fn_l, np.int64),
        rows=np.concatenate(rows_l), row_offsets=np.cumsum([0] + [len(r) for r in rows_l]).astype(np.int64))
    print("wrote %d packets, %d decoded, %d rows" % (len(packets), sum(ok_l), sum(len(r) for r in rows_l)))


if __name__ == "__main__":
    main()
