"""UART TLV packets -> point rows.  The oracle (oracle/tlv_oracle.py) is pinned to the reference's
own decoder: tests/golden/tlv/reference_tlv.npz holds what `ReadIWR14xx.read()` (ReadDataIWR1443.py:27-201, run
unmodified by oracle/gen_tlv_golden.py under a shim that restores the two numpy-1.26 int16 casting behaviours it relies
on) returns for 27 packets over the reference's three radar profiles -- negative coordinates, the Doppler wrap quirk,
peakVal beyond int16, wrong TLV type, no objects, an incomplete packet.  The CUDA decoder is held to those vectors and
to the oracle on arbitrary words."""
import os
import struct

import numpy as np
import pytest

from mmwave_msc_b200 import synth
from oracle import tlv_oracle as tlv

BINS, RES, Q = 32, 0.0626, 9


def _packets(sc):
    return [tlv.encode_packet(f + 1, fr, Q, RES) for f, fr in enumerate(sc.frames)]


def test_oracle_known_vectors():
    # one object, words written by hand: rangeIdx 7, dopplerIdx -3 (0xFFFD), peakVal 40000 (wraps to -25536, the
    # reference stores it in an int16 array), x -1.5 m (= -768 / 512 -> 0xFD00), y 2.25 m, z -0.001953125 m (-1/512)
    body = struct.pack("<6H", 7, 0xFFFD, 40000, 0xFD00, 1152, 0xFFFF)
    t = struct.pack("<2I2H", 1, 4 + len(body), 1, Q) + body
    pkt = tlv.MAGIC + struct.pack("<7I", 0x01020003, 36 + len(t), 0xA1443, 77, 0, 1, 1) + t
    ok, frame, rows = tlv.decode_packet(pkt, BINS, RES)
    assert (ok, frame) == (1, 77)
    np.testing.assert_array_equal(rows, [[-1.5, 2.25, -1 / 512, -3 * RES, -25536.0]])
    # dopplerIdx above numDopplerBins/2 - 1 = 15 as a POSITIVE word: int16 - 65535 wraps to idx + 1 (:167-175)
    body = struct.pack("<6H", 0, 20, 1, 0, 512, 0)
    t = struct.pack("<2I2H", 1, 4 + len(body), 1, Q) + body
    pkt2 = tlv.MAGIC + struct.pack("<7I", 0, 36 + len(t), 0, 5, 0, 1, 1) + t
    assert tlv.decode_packet(pkt2, BINS, RES)[2][0, 3] == 21 * RES
    # no detected objects / other TLV type / bad magic / truncated -> dataOK = 0
    assert tlv.decode_packet(tlv.encode_packet(9, np.zeros((0, 5)), Q, RES), BINS, RES)[0] == 0
    assert tlv.decode_packet(tlv.encode_packet(9, np.ones((3, 5)), Q, RES, tlv_type=2), BINS, RES)[0] == 0
    assert tlv.decode_packet(b"\x00" + pkt[1:], BINS, RES)[0] == 0
    assert tlv.decode_packet(pkt[:-3], BINS, RES)[0] == 0


def _reference_vectors():
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tlv", "reference_tlv.npz"))
    po, ro = g["packet_offsets"], g["row_offsets"]
    for i in range(len(po) - 1):
        yield (bytes(g["packet_bytes"][po[i]:po[i + 1]]), float(g["params"][i, 0]), float(g["params"][i, 1]),
               int(g["ok"][i]), int(g["frame"][i]), g["rows"][ro[i]:ro[i + 1]])


def test_oracle_matches_reference_decoder_vectors():
    n_ok = 0
    for pkt, bins, res, ok, frame, rows in _reference_vectors():
        wok, wframe, wrows = tlv.decode_packet(pkt, bins, res)
        assert (wok, wframe) == (ok, frame)
        if ok:
            np.testing.assert_array_equal(wrows, rows)            # float64, bit for bit
            n_ok += 1
        else:
            assert wrows is None
    assert n_ok == 18


@pytest.mark.gpu
def test_device_decoder_matches_reference_decoder_vectors():
    """Context whose Doppler unit is the profile's dopplerResolutionMps: the rows carry dopplerIdx and
    row[3] * doppler_res in float64 is the reference's doppler value bit for bit; x, y, z, peakVal are exact."""
    from mmwave_msc_b200.batched import BatchedTracker, default_config
    vecs = list(_reference_vectors())
    for res in sorted({v[2] for v in vecs}):
        sel = [v for v in vecs if v[2] == res]
        bt = BatchedTracker(1, config=default_config(doppler_res=res))
        pts, off, frames, ok = bt.decode_tlv([v[0] for v in sel], sel[0][1], res)
        for i, (pkt, bins, _, wok, wframe, rows) in enumerate(sel):
            assert bool(ok[i]) == bool(wok) and int(frames[i]) == wframe
            got = pts[off[i]:off[i + 1]].astype(np.float64)
            if wok:
                np.testing.assert_array_equal(got[:, [0, 1, 2, 4]], rows[:, [0, 1, 2, 4]])
                np.testing.assert_array_equal(got[:, 3] * res, rows[:, 3])
                world, _ = bt.preprocess(pts[off[i]:off[i + 1]])   # ... and on the device
                np.testing.assert_array_equal(world[:, 6], rows[:, 3])
            else:
                assert len(got) == 0
        bt.close()


def test_oracle_round_trip_of_synthetic_frames():
    sc = synth.gen_scene(12, 20)
    for f, pkt in enumerate(_packets(sc)):
        ok, frame, rows = tlv.decode_packet(pkt, BINS, RES)
        assert ok == (1 if len(sc.frames[f]) else 0) and frame == f + 1
        if ok:
            np.testing.assert_array_equal(rows.astype(np.float32), sc.frames[f])


@pytest.mark.gpu
def test_device_decoder_matches_oracle():
    from mmwave_msc_b200.batched import BatchedTracker
    rng = np.random.default_rng(5)
    pkts = []
    for sid in (1, 2, 3):
        pkts += _packets(synth.gen_scene(sid, 12))
    for _ in range(40):                                            # arbitrary 16-bit words, all quirks exercised
        n = int(rng.integers(1, 300))
        body = rng.integers(0, 65536, size=6 * n, dtype=np.uint16).astype("<u2").tobytes()
        t = struct.pack("<2I2H", 1, 4 + len(body), n, int(rng.integers(0, 12))) + body
        pkts.append(tlv.MAGIC + struct.pack("<7I", 0, 36 + len(t), 0, int(rng.integers(0, 1 << 31)), 0, n, 1) + t
                    + bytes(rng.integers(0, 256, size=int(rng.integers(0, 9)), dtype=np.uint8)))
    good = _packets(synth.gen_scene(4, 1))[0]
    pkts += [tlv.encode_packet(3, np.zeros((0, 5)), Q, RES), tlv.encode_packet(4, np.ones((5, 5)), Q, RES, tlv_type=6),
             b"\x09" + good[1:], good[:-5], good[:20], b""]
    bt = BatchedTracker(1)
    pts, off, frames, ok = bt.decode_tlv(pkts, BINS, RES)
    assert off[-1] == len(pts)
    for i, p in enumerate(pkts):
        wok, wframe, rows = tlv.decode_packet(p, BINS, RES)
        assert bool(ok[i]) == bool(wok) and int(frames[i]) == (wframe if wframe < (1 << 31) else wframe - (1 << 32)), i
        got = pts[off[i]:off[i + 1]]
        if wok:
            np.testing.assert_array_equal(got, rows.astype(np.float32), err_msg="packet %d" % i)
        else:
            assert len(got) == 0


@pytest.mark.gpu
def test_packets_to_tracker_equals_frames_to_tracker():
    """Wire bytes -> mmw_decode_tlv -> mmw_step gives the same tracks as stepping with the frames themselves."""
    from mmwave_msc_b200.batched import BatchedTracker
    S, F = 6, 25
    scenes = [synth.gen_scene(s, F) for s in range(20, 20 + S)]
    batches = synth.gen_batch(list(range(20, 20 + S)), F)
    a, b = BatchedTracker(S), BatchedTracker(S)
    for f in range(F):
        pts, off, frames, ok = a.decode_tlv([tlv.encode_packet(f + 1, sc.frames[f], Q, RES) for sc in scenes], BINS, RES)
        np.testing.assert_array_equal(pts, batches[f].points)
        np.testing.assert_array_equal(off, batches[f].offsets)
        a.step(pts, off, batches[f].dt, pose=False)
        b.step(batches[f].points, batches[f].offsets, batches[f].dt, pose=False)
    ta, na = a.tracks()
    tb, nb = b.tracks()
    np.testing.assert_array_equal(na, nb)
    assert na.sum() > 5
    for name in ta.dtype.names:             # value equality: the wire format has no -0.0 Doppler, the generator does
        np.testing.assert_array_equal(ta[name], tb[name], err_msg=name)
