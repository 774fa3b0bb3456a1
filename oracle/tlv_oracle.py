"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's UART TLV decoder for one framed packet
(ReadDataIWR1443.py:88-201, class ReadIWR14xx.read after the magic-word search).

PARITY PINNED (under a numpy-1.x casting shim).  The reference decoder cannot run as it is in this image: it stores
``np.matmul(bytes, [1, 256])`` (0..65535) into ``int16`` arrays and evaluates ``int16_array - 65535`` for every frame
(even when the mask that selects its operand is empty), both of which raise OverflowError under numpy >= 2 (NEP 50;
probed with numpy 2.3.5), and its module imports ``serial``.  ``oracle/gen_tlv_golden.py`` runs ``ReadIWR14xx.read()``
unmodified with a stub ``serial`` and a proxy ``np`` whose int16 arrays restore exactly those two numpy-1.26
behaviours (requirements.txt:8: C-cast wrap on assignment; value-based promotion of ``int16 - 65535`` to int32), over
the reference's own three radar profiles parsed by its own ``__parseConfigFile``.  The 27 packets and what the
reference returned for them are ``tests/golden/tlv/reference_tlv.npz``; ``tests/test_tlv.py`` holds this restatement to
them bit for bit (float64) and the CUDA decoder to both.  What the restatement says:

* header, little endian: magic 8 B | version 4 | totalPacketLen 4 | platform 4 | frameNumber 4 | timeCpuCycles 4 |
  numDetectedObj 4 | numTLVs 4  (ReadDataIWR1443.py:90-98; no subFrameNumber: SDK 1 layout, :101-103)
* only the FIRST TLV is read, and only if numDetectedObj > 0 (:106-107); it must be type 1 =
  MMWDEMO_UART_MSG_DETECTED_POINTS (:115), else the frame yields dataOK = 0
* TLV body: numObj u16, xyzQFormat u16 (:122-129), then per object rangeIdx, dopplerIdx, peakVal, x, y, z, each a
  little-endian 16-bit word stored into an int16 array, i.e. two's complement (:139-163; the explicit
  ``x[x > 32767] -= 65536`` lines are commented out because the cast already does it)
* dopplerIdx[dopplerIdx > numDopplerBins/2 - 1] -= 65535 (:167-175): on int16 data this is ``idx + 1`` (wrapped), a
  quirk that only bites for indices the sensor never sends (it sends them as negative 16-bit numbers already)
* doppler = dopplerIdx * dopplerResolutionMps, x, y, z = int16 / 2**xyzQFormat (:176-184), float64

The device decoder emits float32 rows [x, y, z, doppler, peakVal] -- the tracker's input format; x, y, z, peakVal are
exact; the Doppler column is in units of the context's mmw_config::doppler_res: with doppler_res =
dopplerResolutionMps it is dopplerIdx itself and the device forms the float64 product the reference forms; with the
default unit 1.0 it is that product rounded to float32 (the value the synthetic generator stores).
"""
from __future__ import annotations

import struct
from typing import Optional, Tuple

import numpy as np

MAGIC = bytes([2, 1, 4, 3, 6, 5, 8, 7])
HEADER_BYTES = 36
TLV_DETECTED_POINTS = 1


def decode_packet(pkt: bytes, num_doppler_bins: float, doppler_res: float) -> Tuple[int, int, Optional[np.ndarray]]:
    """(dataOK, frameNumber, rows float64 (numObj, 5) [x, y, z, doppler, peakVal] or None)."""
    if len(pkt) < HEADER_BYTES or pkt[:8] != MAGIC:
        return 0, 0, None
    _ver, total_len, _plat, frame, _cyc, n_det, _n_tlv = struct.unpack_from("<7I", pkt, 8)
    if len(pkt) < total_len:
        return 0, 0, None                              # the reference waits for the rest (:81-83)
    if n_det == 0:
        return 0, int(frame), None
    if len(pkt) < HEADER_BYTES + 12:
        return 0, int(frame), None
    tlv_type, _tlv_len = struct.unpack_from("<2I", pkt, HEADER_BYTES)
    if tlv_type != TLV_DETECTED_POINTS:
        return 0, int(frame), None
    n_obj, q = struct.unpack_from("<2H", pkt, HEADER_BYTES + 8)
    body = HEADER_BYTES + 12
    if len(pkt) < body + 12 * n_obj:
        return 0, int(frame), None
    w = np.frombuffer(pkt, dtype="<i2", count=6 * n_obj, offset=body).reshape(n_obj, 6).astype(np.int16)
    dop_idx = w[:, 1].astype(np.int32)
    over = dop_idx > (num_doppler_bins / 2 - 1)
    dop_idx = np.where(over, (dop_idx - 65535).astype(np.int16).astype(np.int32), dop_idx)     # numpy 1.26 wrap
    scale = float(2 ** int(q))
    rows = np.stack([w[:, 3] / scale, w[:, 4] / scale, w[:, 5] / scale, dop_idx * float(doppler_res),
                     w[:, 2].astype(np.float64)], axis=1)
    return 1, int(frame), rows


def encode_packet(frame_number: int, rows: np.ndarray, q: int, doppler_res: float, tlv_type: int = 1,
                  extra_tail: bytes = b"") -> bytes:
    """Inverse of the above for lattice-valued rows (x, y, z multiples of 2**-q, doppler multiples of doppler_res):
    what an xWR14xx demo would send for this frame (one detected-points TLV)."""
    rows = np.asarray(rows, np.float64).reshape(-1, 5)
    n = rows.shape[0]
    xyz = np.round(rows[:, :3] * (1 << q)).astype(np.int64)
    dop = np.round(rows[:, 3] / doppler_res).astype(np.int64)
    peak = np.round(rows[:, 4]).astype(np.int64)
    rng = np.round(np.sqrt((rows[:, :3] ** 2).sum(1)) / 0.047).astype(np.int64)        # rangeIdx: unused downstream
    words = np.stack([rng, dop, peak, xyz[:, 0], xyz[:, 1], xyz[:, 2]], axis=1)
    body = (words & 0xffff).astype("<u2").tobytes()
    tlv = struct.pack("<2I2H", tlv_type, 4 + len(body), n, q) + body
    total = HEADER_BYTES + len(tlv) + len(extra_tail)
    hdr = MAGIC + struct.pack("<7I", 0x01020003, total, 0xA1443, frame_number, 0, n, 1 if n else 0)
    return hdr + (tlv if n else b"") + extra_tail
