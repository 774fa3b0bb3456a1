"""The grid screen in front of DBSCAN in the tracker step (mmwave_msc_b200/csrc/dbscan.cuh: grid_screen_config,
grid_cell, grid_screen_may_have_core; the per-frame histograms are filled in step_kernel.cu when a frame is pushed),
restated in numpy with the kernel's fp32 arithmetic and held against the oracle's exact eps-neighbourhoods
(Utils.py:222-247 through oracle.mmw_oracle.pair_distance_matrix).

What must hold for the CUDA path to stay bit-exact: whenever the screen answers "no point can be a core point", no
point of the cloud has min_samples neighbours -- for clouds inside the grid, outside it, beyond the range bound the
cells were sized for, and where the range weight turns negative.  What makes it worth having: on the residue clouds of
steady scenes it answers "no" almost always.  (The device code itself is compared with the oracle in
test_gpu_parity.py::test_grid_screen_edge_cases and by every sequence test.)"""
import numpy as np

from mmwave_msc_b200 import synth
from oracle import mmw_oracle as mo

GRID = 16


def grid_config(cfg: mo.OracleConfig):
    """grid_screen_config: (inv_h, ybound, ok) in fp32 like the host code."""
    f32 = np.float32
    if not (cfg.db_range_weight >= 0.0) or not (cfg.db_z_weight >= 0.0) or not (cfg.db_eps > 0.0) \
            or cfg.db_min_samples > 255:
        return f32(0), f32(0), False
    yb = 12.0
    if cfg.db_range_weight * yb > 0.5:
        yb = 0.5 / cfg.db_range_weight
    wmin = f32(1) - f32(cfg.db_range_weight) * (f32(yb) * f32(1.0001))
    if not (wmin > f32(0.05)):
        return f32(0), f32(0), False
    h = f32(1.01) * np.sqrt(f32(cfg.db_eps) / wmin, dtype=f32)
    return f32(1) / h, f32(yb), True


def screen_may_have_core(world: np.ndarray, cfg: mo.OracleConfig) -> bool:
    """The step kernel's screen: world = (B, >=2) float64 world-frame points (x, y', ...) of the fused ring."""
    B = len(world)
    if B < cfg.db_min_samples:
        return False
    inv_h, ybound, ok = grid_config(cfg)
    if not ok:
        return True
    f32 = np.float32
    X, Y = world[:, 0].astype(f32), world[:, 1].astype(f32)          # raw x is fp32; y' is rounded to fp32 for the cell
    if not np.all(Y <= ybound):                                       # the frame's flag byte: the screen passes
        return True
    cx = np.clip(np.floor(X * inv_h + f32(0.5 * GRID)).astype(np.int64), 0, GRID - 1)
    cy = np.clip(np.floor(Y * inv_h).astype(np.int64), 0, GRID - 1)
    hist = np.zeros((GRID + 2, GRID + 2), np.int64)                   # one ring of empty cells instead of bounds checks
    np.add.at(hist, (cy + 1, cx + 1), 1)
    hist = np.minimum(hist, 3 * 255)                                  # three frames of saturating byte counts
    block = sum(hist[cy + 1 + dy, cx + 1 + dx] for dy in (-1, 0, 1) for dx in (-1, 0, 1))
    return bool((block >= cfg.db_min_samples).any())


def has_core_point(world: np.ndarray, cfg: mo.OracleConfig) -> bool:
    if len(world) == 0:
        return False
    adj = mo.pair_distance_matrix(np.asarray(world, np.float64)[:, :3], cfg) <= cfg.db_eps
    return bool((adj.sum(axis=1) >= cfg.db_min_samples).any())


def _residue_clouds(scene_ids, n_frames, skip):
    """Fused unassigned clouds that TrackBuffer.track hands to DBSCAN (Tracking.py:693-697) in oracle-run scenes."""
    cfg = mo.OracleConfig()
    clouds = []
    orig = mo.dbscan_labels

    def capture(points, *a, **k):
        clouds.append(np.array(points, np.float64))
        return orig(points, *a, **k)

    mo.dbscan_labels = capture
    try:
        for sid in scene_ids:
            sc = synth.gen_scene(sid, n_frames)
            so = mo.SceneOracle(cfg)
            for f, (raw, dt) in enumerate(zip(sc.frames, sc.dts())):
                n0 = len(clouds)
                so.step(raw, float(dt))
                if f < skip:
                    del clouds[n0:]
    finally:
        mo.dbscan_labels = orig
    return cfg, clouds


def test_screen_never_hides_a_core_point_on_scene_residues_and_skips_most():
    cfg, clouds = _residue_clouds(range(40, 50), 36, skip=0)          # start-up clouds (people walk in) included
    assert len(clouds) > 200
    with_core = 0
    for c in clouds:
        core = has_core_point(c, cfg)
        with_core += core
        if core:
            assert screen_may_have_core(c, cfg)
    assert with_core >= 5                                             # the implication was exercised
    cfg, steady = _residue_clouds(range(50, 58), 40, skip=14)
    skipped = sum(not screen_may_have_core(c, cfg) for c in steady)
    assert skipped >= 0.9 * len(steady), "the screen should dismiss the residue of a steady scene (%d of %d)" % (
        skipped, len(steady))


def test_screen_is_conservative_on_adversarial_clouds():
    """Blobs anywhere (also beyond the 16 x 16 cells and far down range), sized around min_samples, packed so that
    neighbours straddle cell borders; with and without clutter."""
    cfg = mo.OracleConfig()
    rng = np.random.default_rng(20240)
    reach = float(np.sqrt(cfg.db_eps))
    n_core = n_screened_out = 0
    for trial in range(400):
        parts = []
        for _ in range(int(rng.integers(1, 4))):
            n = int(rng.integers(cfg.db_min_samples - 12, cfg.db_min_samples + 25))
            cx, cy = rng.uniform(-8.0, 8.0), rng.uniform(0.2, 32.0 if trial % 4 else 45.0)
            s = rng.choice([0.02, 0.1, 0.25, 0.5]) * reach
            parts.append(np.stack([rng.normal(cx, s, n), np.abs(rng.normal(cy, s, n)) + 1e-3, rng.uniform(0.1, 2.4, n)], 1))
        if trial % 2:
            m = int(rng.integers(10, 120))
            parts.append(np.stack([rng.uniform(-2.5, 2.5, m), rng.uniform(0.5, 8.5, m), rng.uniform(0.1, 2.4, m)], 1))
        cloud = np.concatenate(parts)
        # the sensor lattice: values as the device sees them (fp32 inputs, up-cast exactly)
        cloud = cloud.astype(np.float32).astype(np.float64)
        core, may = has_core_point(cloud, cfg), screen_may_have_core(cloud, cfg)
        n_core += core
        n_screened_out += not may
        assert may or not core, "trial %d: the screen dismissed a cloud with a core point" % trial
    assert n_core > 100 and n_screened_out > 20                       # both outcomes occurred


def test_screen_degenerate_configurations_fall_back():
    cfg = mo.OracleConfig()
    pts = np.zeros((40, 3)); pts[:, 1] = 1.0
    assert screen_may_have_core(pts, cfg)                             # 40 coincident points
    assert not screen_may_have_core(pts[:cfg.db_min_samples - 1], cfg)
    far = pts.copy(); far[:, 1] = 34.0                                # beyond the range bound (1 - 0.03 * 34 < 0: no reach bound)
    far[:, 0] = np.linspace(-50, 50, 40)
    assert screen_may_have_core(far, cfg)
    assert has_core_point(far, cfg)                                   # ... and indeed: with a negative weight all are neighbours
