"""CPU-side checks of the boundary: the shared library loads, exports every symbol include/mmw.h declares,
struct layouts agree, and the product fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from mmwave_msc_b200 import _lib


def _header_symbols():
    h = open(os.path.join(ROOT, "include", "mmw.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(mmw_[a-z_0-9]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), "libmmw.so does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes SIGNATURES and include/mmw.h disagree"
    want = int(re.search(r"#define MMW_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "mmw.h")).read()).group(1))
    assert lib.mmw_abi_version() == want == _lib.ABI_VERSION


def test_default_config_matches_reference_constants():
    import json
    ka = json.load(open(os.path.join(ROOT, "tests", "golden", "known_answers.json")))["constants"]
    from mmwave_msc_b200.batched import default_config
    c = default_config()
    assert (c.s_height, c.s_tilt_deg, c.frames_batch) == (ka["S_HEIGHT"], ka["S_TILT"], ka["FB_FRAMES_BATCH"])
    assert (c.db_z_weight, c.db_range_weight, c.db_eps, c.db_min_samples) == (
        ka["DB_Z_WEIGHT"], ka["DB_RANGE_WEIGHT"], ka["DB_EPS"], ka["DB_MIN_SAMPLES_MIN"])
    assert (c.tr_max_tracks, c.tr_lifetime_dynamic, c.tr_lifetime_static, c.tr_vel_thres, c.tr_gate) == (
        ka["TR_MAX_TRACKS"], ka["TR_LIFETIME_DYNAMIC"], ka["TR_LIFETIME_STATIC"], ka["TR_VEL_THRES"], ka["TR_GATE"])
    assert (c.kf_q_var, c.kf_p_init, c.kf_group_disp_init, bool(c.kf_enable_est), c.kf_a_n, c.kf_est_pointnum,
            c.kf_a_spr) == (ka["KF_Q_STD"], ka["KF_P_INIT"], ka["KF_GROUP_DISP_EST_INIT"], ka["KF_ENABLE_EST"],
                            ka["KF_A_N"], ka["KF_EST_POINTNUM"], ka["KF_A_SPR"])
    assert list(c.kf_spread_lim) == ka["KF_SPREAD_LIM"]
    assert (c.intensity_mu, c.intensity_std) == (ka["INTENSITY_MU"], ka["INTENSITY_STD"])
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "known_answers.json")))["default_posture"]
    np.testing.assert_allclose(list(c.default_posture), golden, rtol=0, atol=1e-7)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mmwave_msc_b200.batched import BatchedTracker
    with pytest.raises(_lib.MmwError):
        BatchedTracker(1)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mmwave_msc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_offline_manager_matches_reference_reader(tmp_path):
    """The CSV replay reader hands out the same frames as the reference's OfflineManager did on the same log
    (golden sequence recorded by oracle/gen_golden.py), including the dropped rows of every 40th frame."""
    import json
    import sys
    import types
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "offline_manager_sequence.json")))
    from mmwave_msc_b200 import synth
    sc = synth.gen_scene(g["scene"], g["frames"])
    synth.write_reference_csv(sc, str(tmp_path), frames_per_file=g["frames_per_file"])
    # import Utils without touching the device: OfflineManager is pure host code
    from mmwave_msc_b200 import Utils
    om = Utils.OfflineManager(str(tmp_path))
    seq = []
    while not om.is_finished() and len(seq) < 400:
        ok, fc, det = om.get_data()
        seq.append([int(ok), int(fc), len(det["x"]) if ok else -1, int(det["posix"][0]) if ok else -1])
    assert seq == g["sequence"]
