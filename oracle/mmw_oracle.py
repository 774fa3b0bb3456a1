"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float64) of the reference's
per-frame radar perception path.  It is the checker for the CUDA path; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may
import it.  The product (``mmwave_msc_b200``) never does.

Pinning: this restatement is checked against the reference's own modules run in
the build container (``oracle/ref_harness.py`` -> ``tests/golden/*.npz``, see
``oracle/gen_golden.py``) and against the literal golden values of SURVEY.md
section 8(c).  Pose parity against the author's trained network is UNPINNED (the
weights ``MARS.h5`` are not shipped, constants.py:14): the CNN here restates the
architecture of train.py:33-106 and is compared with seeded synthetic weights.

Every function cites the reference file:line it follows (paths relative to
/root/reference/src).  Reference arithmetic is float64 (numpy); so is this.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np


# ----------------------------------------------------------------------------
# configuration (constants.py)
# ----------------------------------------------------------------------------
@dataclass
class OracleConfig:
    s_height: float = 1.8                    # constants.py:41
    s_tilt_deg: float = -5.0                 # constants.py:42
    z_max: float = 2.5                       # Utils.py:424
    frames_batch: int = 2                    # FB_FRAMES_BATCH constants.py:66 (ring holds frames_batch+1)
    db_z_weight: float = 0.4                 # constants.py:70
    db_range_weight: float = 0.03            # constants.py:71
    db_eps: float = 0.3                      # constants.py:72
    db_min_samples: int = 35                 # constants.py:73
    tr_max_tracks: int = 4                   # constants.py:85
    tr_lifetime_dynamic: float = 3.0         # constants.py:86
    tr_lifetime_static: float = 7.0          # constants.py:87
    tr_vel_thres: float = 0.12               # constants.py:88
    tr_gate: float = 4.5                     # constants.py:89
    kf_q_var: float = 1.0                    # KF_Q_STD passed as var, constants.py:93,212
    kf_p_init: float = 0.1                   # constants.py:96
    kf_group_disp_init: float = 0.1          # constants.py:97
    kf_enable_est: bool = False              # constants.py:100
    kf_a_n: float = 0.9                      # constants.py:101
    kf_est_pointnum: int = 10                # constants.py:102
    kf_spread_lim: Sequence[float] = (0.2, 0.2, 2, 1.2, 1.2, 0.2)   # constants.py:103
    kf_a_spr: float = 0.9                    # constants.py:104
    intensity_mu: float = 27.0187            # constants.py:108
    intensity_std: float = 70.351            # constants.py:109
    x_nudge_thres: float = 0.6               # Tracking.py:397
    x_nudge_gain: float = 0.4                # Tracking.py:398


# constants.py:112-172
MODEL_DEFAULT_POSTURE = np.array([
    0.0000, -0.0007, -0.0006, -0.0038, -0.1820, -0.2540, -0.2579, 0.1830, 0.2957, 0.2940,
    -0.0805, -0.1141, -0.1232, -0.1358, 0.0796, 0.1436, 0.1558, 0.1720, -0.0007, 0.7699,
    1.0906, 1.4020, 1.5513, 1.2893, 1.0360, 0.7994, 1.2865, 1.0483, 0.8117, 0.7670,
    0.3428, 0.0000, -0.0746, 0.7713, 0.3706, -0.0128, -0.0796, 1.3255, 0.0752, 0.0533,
    0.0203, 0.0000, 0.0496, 0.1350, 0.1303, 0.0345, 0.1277, 0.1050, 0.0392, 0.0533,
    0.0786, -0.0056, 0.0346, -0.0007, 0.0683, -0.0082, 0.0312])


# ----------------------------------------------------------------------------
# a1  preprocessing  (Utils.py:294-434)
# ----------------------------------------------------------------------------
def normalize_points(raw: np.ndarray, cfg: OracleConfig = OracleConfig()):
    """Utils.normalize_data (Utils.py:342-434) on an (N,5) array x,y,z,doppler,peakVal.

    Returns (world (M,8) float64, keep mask (N,) bool).  Per point:
    r = sqrt(x^2+y^2+z^2) (Utils.py:382-386); r == 0 -> v = (0, doppler, 0)
    (387-390) else v = (doppler*coord)/r (400-402); rotation about X by
    radians(S_TILT) plus S_HEIGHT on z, velocities rotated only
    (point_transform_to_standard_axis, 294-339); keep iff z' <= 2.5 and z' > 0
    and y' > 0 (423-427); order preserved.
    """
    raw = np.asarray(raw, dtype=np.float64).reshape(-1, 5)
    x, y, z, dop, peak = (raw[:, k] for k in range(5))
    r = np.sqrt(x * x + y * y + z * z)
    zero = r == 0
    rs = np.where(zero, 1.0, r)
    vx = np.where(zero, 0.0, dop * x / rs)
    vy = np.where(zero, dop, dop * y / rs)
    vz = np.where(zero, 0.0, dop * z / rs)
    ang = np.radians(cfg.s_tilt_deg)
    c, s = np.cos(ang), np.sin(ang)
    yw = c * y + (-s) * z
    zw = (s * y + c * z) + cfg.s_height
    vyw = c * vy + (-s) * vz
    vzw = s * vy + c * vz
    world = np.stack([x, yw, zw, vx, vyw, vzw, dop, peak], axis=1)
    keep = (zw <= cfg.z_max) & (zw > 0) & (yw > 0)
    return world[keep], keep


# ----------------------------------------------------------------------------
# a2/a3  clustering  (Utils.py:222-291 + sklearn DBSCAN semantics)
# ----------------------------------------------------------------------------
def pair_distance_matrix(xyz: np.ndarray, cfg: OracleConfig = OracleConfig()) -> np.ndarray:
    """altered_EuclideanDist for all pairs (Utils.py:242-247): a weighted SQUARED
    distance, weight = 1 - ((y1+y2)/2)*DB_RANGE_WEIGHT."""
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    w = 1 - ((y[:, None] + y[None, :]) / 2) * cfg.db_range_weight
    dx = x[:, None] - x[None, :]
    dy = y[:, None] - y[None, :]
    dz = z[:, None] - z[None, :]
    return w * (dx ** 2 + dy ** 2 + cfg.db_z_weight * (dz ** 2))


def dbscan_labels(points: np.ndarray, cfg: OracleConfig = OracleConfig(),
                  eps: Optional[float] = None, min_samples: Optional[int] = None) -> np.ndarray:
    """Exact-neighbourhood DBSCAN with sklearn's labelling (SURVEY.md Appendix A
    Q1-Q5; sklearn ``DBSCAN.fit`` + ``dbscan_inner``): neighbourhood =
    {q : d(p,q) <= eps} including p; core iff |neighbourhood| >= min_samples;
    clusters numbered by ascending smallest core index; a border point gets the
    lowest-numbered cluster with a core point within eps; noise = -1."""
    eps = cfg.db_eps if eps is None else eps
    min_samples = cfg.db_min_samples if min_samples is None else min_samples
    n = len(points)
    labels = np.full(n, -1, dtype=np.int32)
    if n == 0:
        return labels
    adj = pair_distance_matrix(np.asarray(points, dtype=np.float64)[:, :3], cfg) <= eps
    core = adj.sum(axis=1) >= min_samples
    lab = 0
    for i in range(n):                      # dbscan_inner: outer loop ascending, DFS with a stack
        if labels[i] != -1 or not core[i]:
            continue
        stack = [i]
        while stack:
            k = stack.pop()
            if labels[k] == -1:
                labels[k] = lab
                if core[k]:
                    for v in np.nonzero(adj[k])[0]:
                        if labels[v] == -1:
                            stack.append(int(v))
        lab += 1
    return labels


def clusters_from_labels(points: np.ndarray, labels: np.ndarray) -> List[np.ndarray]:
    """apply_DBscan's return value (Utils.py:281-291): clusters in ascending label
    order, points in input order, noise dropped."""
    return [points[labels == c] for c in range(int(labels.max()) + 1)] if len(labels) else []


# ----------------------------------------------------------------------------
# a6  motion model (constants.py:176-215) and filterpy semantics (Appendix B)
# ----------------------------------------------------------------------------
def kf_F(dt: float) -> np.ndarray:
    """CONST_ACC_MODEL.KF_F (constants.py:195-208)."""
    F = np.eye(9)
    h = 0.5 * dt ** 2
    for i in range(3):
        F[i, i + 3] = dt
        F[i, i + 6] = h
        F[i + 3, i + 6] = dt
    return F


def kf_Q(dt: float, var: float = 1.0) -> np.ndarray:
    """CONST_ACC_MODEL.KF_Q_DISCR (constants.py:210-215): block_diag of three
    Q_discrete_white_noise(dim=3) blocks -- on state indices (0,1,2), (3,4,5),
    (6,7,8), which is NOT the state's [pos, vel, acc]-per-axis order (Q10)."""
    Qw = np.array([[0.25 * dt ** 4, 0.5 * dt ** 3, 0.5 * dt ** 2],
                   [0.5 * dt ** 3, dt ** 2, dt],
                   [0.5 * dt ** 2, dt, 1.0]]) * var
    Q = np.zeros((9, 9))
    for b in range(3):
        Q[3 * b:3 * b + 3, 3 * b:3 * b + 3] = Qw
    return Q


class Ring:
    """BatchedData / RingBuffer (Tracking.py:21-71, Utils.py:10-50): at most
    ``size`` frames, oldest dropped first; fused view = concatenation oldest ->
    newest."""

    def __init__(self, size: int, init: Optional[np.ndarray] = None):
        self.size = size
        self.frames: List[np.ndarray] = [np.empty((0, 8)) if init is None else init]

    def add(self, frame: np.ndarray):
        while len(self.frames) >= self.size:        # Tracking.py:47-48
            self.frames.pop(0)
        self.frames.append(frame)

    def pop(self):
        if self.frames:
            self.frames.pop(0)

    def clear(self):
        self.frames = []

    def fused(self) -> np.ndarray:
        return np.concatenate(self.frames, axis=0) if self.frames else np.empty((0, 8))


class Track:
    """ClusterTrack + KalmanState + PointCluster (Tracking.py:74-407)."""

    def __init__(self, cloud: np.ndarray, track_id: int, cfg: OracleConfig):
        self.cfg = cfg
        self.id = track_id
        self.N_est = 0.0                                        # Tracking.py:211
        self.spread_est = np.zeros(6)                            # :212
        self.group_disp_est = np.eye(6) * cfg.kf_group_disp_init  # :213-215
        self._set_cluster(cloud)
        self.ring = Ring(cfg.frames_batch + 1, cloud)            # :217
        self.x = np.concatenate([self.centroid, np.zeros(3)])    # :96, constants.py:191-192
        self.P = np.eye(9) * cfg.kf_p_init                       # :97
        self.lifetime = 0.0
        self.keypoints = MODEL_DEFAULT_POSTURE.copy()            # :221

    def _set_cluster(self, cloud: np.ndarray):
        """PointCluster.__init__ (Tracking.py:120-136)."""
        self.cloud = cloud
        self.point_num = cloud.shape[0]
        self.centroid = np.mean(cloud[:, :6], axis=0)
        self.min_vals = np.min(cloud[:, :6], axis=0)
        self.max_vals = np.max(cloud[:, :6], axis=0)
        self.static = math.sqrt(np.sum(self.centroid[3:6] ** 2)) < self.cfg.tr_vel_thres

    def predict(self, dt: float):
        """ClusterTrack.predict_state (Tracking.py:372-385) -> filterpy predict."""
        F, Q = kf_F(dt), kf_Q(dt, self.cfg.kf_q_var)
        self.x = F @ self.x
        self.P = (F @ self.P) @ F.T + Q

    def associate(self, cloud: np.ndarray):
        """ClusterTrack.associate_pointcloud (Tracking.py:314-341)."""
        cfg = self.cfg
        self._set_cluster(cloud)
        self.ring.add(cloud)
        N = self.point_num
        if cfg.kf_enable_est:                                    # :236-242
            if N > self.N_est:
                self.N_est = float(N)
            else:
                self.N_est = (1 - cfg.kf_a_n) * self.N_est + cfg.kf_a_n * N
        else:
            self.N_est = float(max(cfg.kf_est_pointnum, N))      # :244
        for m in range(6):                                       # :250-268
            spread = self.max_vals[m] - self.min_vals[m]
            if N != 1:
                spread = spread * (N + 1) / (N - 1)
            spread = min(2 * cfg.kf_spread_lim[m], spread)
            spread = max(cfg.kf_spread_lim[m], spread)
            if spread > self.spread_est[m]:
                self.spread_est[m] = spread
            else:
                self.spread_est[m] = (1.0 - cfg.kf_a_spr) * self.spread_est[m] + cfg.kf_a_spr * spread
        dev = cloud[:, :6] - self.centroid                       # _get_D :270-290 (population covariance)
        D = (dev[:, :, None] * dev[:, None, :]).mean(axis=0)
        a = N / self.N_est                                       # :296-297
        self.group_disp_est = (1 - a) * self.group_disp_est + a * D

    def Rm(self) -> np.ndarray:
        return np.diag((self.spread_est / 2) ** 2)               # get_Rm :361-370

    def Rc(self) -> np.ndarray:
        N, N_est = self.point_num, self.N_est                    # _get_Rc :299-312
        return (self.Rm() / N) + ((N_est - N) / ((N_est - 1) * N)) * self.group_disp_est

    def update(self):
        """ClusterTrack.update_state (Tracking.py:387-398) -> filterpy Joseph update,
        then the x[0] nudge of Q12."""
        z = self.centroid
        R = self.Rc()
        y = z - self.x[:6]
        PHT = self.P[:, :6]
        S = PHT[:6, :] + R
        SI = np.linalg.inv(S)
        K = PHT @ SI
        self.x = self.x + K @ y
        I_KH = np.eye(9)
        I_KH[:, :6] -= K
        self.P = (I_KH @ self.P) @ I_KH.T + (K @ R) @ K.T
        variance = z[0] - self.x[0]
        if abs(bool(variance != 0)) > self.cfg.x_nudge_thres and self.lifetime == 0:   # :397 (Q12)
            self.x[0] += variance * self.cfg.x_nudge_gain


def gate_scores(world: np.ndarray, tracks: Sequence[Track]) -> np.ndarray:
    """d^2 matrix of TrackBuffer._calc_dist_fun (Tracking.py:545-560):
    log|det C| + y' C^-1 y with C = P[:6,:6] + Rm + group_disp_est."""
    d2 = np.empty((world.shape[0], len(tracks)))
    for j, tr in enumerate(tracks):
        C = tr.P[:6, :6] + tr.Rm() + tr.group_disp_est
        Ci = np.linalg.inv(C)
        logdet = np.log(np.abs(np.linalg.det(C)))
        Y = world[:, :6] - tr.x[:6]
        d2[:, j] = logdet + ((Y @ Ci) * Y).sum(axis=1)
    return d2


def associate_from_scores(d2: np.ndarray, gate: float) -> np.ndarray:
    """Gate (strict <) and strict-< arg-min so ties go to the lower track index
    (Tracking.py:563-572).  -1 = unassigned (None in the reference)."""
    n, T = d2.shape
    assoc = np.full(n, -1, dtype=np.int32)
    best = np.full(n, np.inf)
    for j in range(T):
        better = (d2[:, j] < gate) & (d2[:, j] < best)
        assoc[better] = j
        best[better] = d2[better, j]
    return assoc


# ----------------------------------------------------------------------------
# a12/a13  pose features  (Utils.py:437-520, Tracking.py:720-728)
# ----------------------------------------------------------------------------
def pose_features(track: Track, cfg: OracleConfig = OracleConfig()) -> np.ndarray:
    """relative_coordinates + format_single_frame with the canonical (stable)
    tie order of Q21.  Returns (frames_batch+1, 8, 8, 5), or (8, 8, 5) when
    frames_batch == 0 (Utils.py:517-520)."""
    nfr = cfg.frames_batch + 1
    out = np.zeros((nfr, 64, 5))
    cx, cy = track.centroid[0], track.centroid[1]               # current centroid for every frame (722-725)
    for f, cloud in enumerate(track.ring.frames):
        rel = cloud - np.array([cx, cy, 0, 0, 0, 0, 0, 0])       # Utils.py:454-463
        sel = rel[:, [0, 1, 2, 6, 7]].copy()                     # :498
        sel[:, 4] = (sel[:, 4] - cfg.intensity_mu) / cfg.intensity_std   # :502
        if len(sel) < 64:                                        # :505-510
            sel = np.concatenate([sel, np.zeros((64 - len(sel), 5))], axis=0)
        else:
            sel = sel[:64]
        out[f] = sel[np.argsort(sel[:, 0], kind="stable")]       # :513-514 with canonical ties
    if cfg.frames_batch == 0:
        return out.reshape(8, 8, 5)
    return out.reshape(nfr, 8, 8, 5)


# ----------------------------------------------------------------------------
# a15/a16  pose CNN  (train.py:33-106), inference semantics of Keras
# ----------------------------------------------------------------------------
def pose_forward(weights: Sequence[np.ndarray], feats: np.ndarray, dtype=np.float64) -> np.ndarray:
    """define_CNN (2-D, feats (n,8,8,5)) / define_CNN_3D (feats (n,3,8,8,5)).

    ``weights`` is Keras ``model.get_weights()`` order: conv1 kernel
    (k..,cin,cout), bias, conv2 kernel, bias, BN1 gamma, beta, moving_mean,
    moving_var, dense1 kernel (in,out), bias, BN2 (4), dense2 kernel, bias.
    Conv = cross-correlation, zero padding 'same', ReLU; Dropout = identity;
    BatchNorm with eps 1e-3 on the channel axis; Flatten in channels-last order.
    """
    import torch
    import torch.nn.functional as Fn
    td = torch.float64 if dtype == np.float64 else torch.float32
    w = [torch.from_numpy(np.asarray(a)).to(td) for a in weights]
    x = torch.from_numpy(np.asarray(feats)).to(td)
    eps = 1e-3
    three_d = x.dim() == 5
    with torch.no_grad():
        if three_d:
            x = x.permute(0, 4, 1, 2, 3)                         # NDHWC -> NCDHW
            k1 = w[0].permute(4, 3, 0, 1, 2)
            k2 = w[2].permute(4, 3, 0, 1, 2)
            h = torch.relu(Fn.conv3d(x, k1, w[1], padding=1))
            h = torch.relu(Fn.conv3d(h, k2, w[3], padding=1))
            h = h.permute(0, 2, 3, 4, 1)                         # channels last
        else:
            x = x.permute(0, 3, 1, 2)
            k1 = w[0].permute(3, 2, 0, 1)
            k2 = w[2].permute(3, 2, 0, 1)
            h = torch.relu(Fn.conv2d(x, k1, w[1], padding=1))
            h = torch.relu(Fn.conv2d(h, k2, w[3], padding=1))
            h = h.permute(0, 2, 3, 1)
        h = (h - w[6]) / torch.sqrt(w[7] + eps) * w[4] + w[5]
        h = h.reshape(h.shape[0], -1)
        h = torch.relu(h @ w[8] + w[9])
        h = (h - w[12]) / torch.sqrt(w[13] + eps) * w[10] + w[11]
        out = h @ w[14] + w[15]
    return out.numpy()


# ----------------------------------------------------------------------------
# "next" row f-2: smart-window projection of a track (Utils.py:180-219, Visualizer.py:14-29)
# ----------------------------------------------------------------------------
def projection_point(x_origin, y_origin, z_origin, m=(0.32, -0.6, 1.3)):
    """Utils.calc_projection_points (Utils.py:180-219) with M_X, M_Y, M_Z of constants.py:31-33."""
    x_dist, y_dist, z_dist = x_origin - m[0], y_origin - m[1], z_origin - m[2]
    x_proj = x_origin if x_dist == 0 else -m[1] / (y_dist / x_dist) + m[0]
    z_proj = z_origin if z_dist == 0 else -m[1] / (y_dist / z_dist) + m[2]
    return x_proj, z_proj


def fade_square(x, keypoints, m=(0.32, -0.6, 1.3), size_max=0.3, size_min=0.2, weight=0.08):
    """Visualizer.calc_fade_square (Visualizer.py:14-29): centre and side of the opaque square.  Restated from the
    source (Visualizer.py cannot be imported here: matplotlib / pyqtgraph are absent), so only its
    calc_projection_points part is pinned by a golden value."""
    kp = np.asarray(keypoints, dtype=np.float64)
    center = projection_point(x[0] + kp[3], x[1] + kp[41], kp[22], m)
    size = max(size_min, min(size_max, size_max - (x[1] + kp[12]) * weight))
    return center, size


# ----------------------------------------------------------------------------
# "next" row f-3: dataset-builder export of track 0 (preprocessing.py:185-216, Utils.py:523-548)
# ----------------------------------------------------------------------------
def track0_export(scene: "SceneOracle"):
    """What preprocess_dataset writes for a frame that ran track(): None unless track 0 exists, was seen in this
    frame (lifetime == 0) and has points in its ring (preprocessing.py:185-196); else the (192, 5) block of
    format_batched_frames(relative_coordinates(ring frames, centroid)) -- ring frames newest first, columns
    [x - cx, y - cy, z, doppler, peakVal], first 64 rows per frame, zero padded, unsorted, raw intensity
    (Utils.py:523-548, 437-465) -- and the centroid[:2] appended to the centroid log (preprocessing.py:209-213)."""
    if not scene.tracks:
        return None
    t0 = scene.tracks[0]
    if t0.lifetime != 0 or sum(len(f) for f in t0.ring.frames) == 0:
        return None
    out = np.zeros((3 * 64, 5))
    cx, cy = t0.centroid[0], t0.centroid[1]
    for k, cloud in enumerate(reversed(t0.ring.frames)):
        rel = cloud - np.array([cx, cy, 0, 0, 0, 0, 0, 0])
        sel = rel[:, [0, 1, 2, 6, 7]][:64]
        out[k * 64:k * 64 + len(sel)] = sel
    return out, np.array([cx, cy])


# ----------------------------------------------------------------------------
# a11  one scene: TrackBuffer.track + estimate_posture under the offline_main loop
# ----------------------------------------------------------------------------
class SceneOracle:
    """State of one scene: TrackBuffer + the global BatchedData ring
    (offline_main.py:32-34) stepped by the loop body of offline_main.py:45-60."""

    def __init__(self, cfg: OracleConfig = OracleConfig(), pose_weights=None, pose_dtype=np.float64):
        self.cfg = cfg
        self.tracks: List[Track] = []
        self.next_track_id = 0
        self.ring = Ring(cfg.frames_batch + 1)
        self.pose_weights = pose_weights
        self.pose_dtype = pose_dtype

    def step(self, raw: np.ndarray, dt: float) -> dict:
        cfg = self.cfg
        world, keep = normalize_points(raw, cfg)
        rec = {"M": int(world.shape[0]), "world": world, "keep": keep, "ran": world.shape[0] != 0,
               "assoc": np.zeros(0, np.int32), "labels": None, "features": None, "d2": None}
        if world.shape[0] != 0:                                  # offline_main.py:55 (Q23)
            self._track(world, float(dt), rec)
            self._estimate_posture(rec)
        rec["ring_counts"] = np.array([len(f) for f in self.ring.frames], dtype=np.int32)
        rec["next_track_id"] = self.next_track_id
        rec["tracks"] = [self._snapshot(t) for t in self.tracks]
        return rec

    def _track(self, world: np.ndarray, dt: float, rec: dict):
        """TrackBuffer.track (Tracking.py:664-703)."""
        cfg = self.cfg
        for tr in self.tracks:                                   # _predict_all :591-596 (Q9)
            tr.predict(tr.lifetime + dt)
        d2 = gate_scores(world, self.tracks)                     # _calc_dist_fun :530-574
        assoc = associate_from_scores(d2, cfg.tr_gate)
        rec["assoc"], rec["d2"] = assoc, d2
        for j, tr in enumerate(self.tracks):                     # _associate_points_to_tracks :648-653
            cloud = world[assoc == j]
            if len(cloud) == 0:
                tr.lifetime += dt
            else:
                tr.lifetime = 0.0
                tr.associate(cloud)
        self.tracks = [tr for tr in self.tracks if not (          # _maintain_tracks :513-528 (Q16)
            tr.lifetime > (cfg.tr_lifetime_static if tr.static else cfg.tr_lifetime_dynamic))]
        for tr in self.tracks:                                   # _update_all :598-603 (Q11)
            tr.update()
        self.ring.add(world[assoc == -1])                        # :691 (Q22)
        fused = self.ring.fused()
        if len(fused) > 0 and len(self.tracks) < cfg.tr_max_tracks:   # :693-696 (Q7)
            labels = dbscan_labels(fused, cfg)                   # apply_DBscan :697
            rec["labels"] = labels
            clusters = clusters_from_labels(fused, labels)
            if len(clusters) > 0:                                # :699-700 (Q8)
                self.ring.clear()
            for c in clusters:                                   # _add_tracks :576-589
                self.tracks.append(Track(c, self.next_track_id, cfg))
                self.next_track_id += 1

    def _estimate_posture(self, rec: dict):
        """TrackBuffer.estimate_posture (Tracking.py:705-734)."""
        feats, idx = [], []
        for i, tr in enumerate(self.tracks):
            if sum(len(f) for f in tr.ring.frames) > 0:          # MODEL_MIN_INPUT = 0 (:721)
                feats.append(pose_features(tr, self.cfg))
                idx.append(i)
        if not feats:
            return
        feats = np.array(feats)
        rec["features"] = feats
        if self.pose_weights is not None:
            kp = pose_forward(self.pose_weights, feats.astype(np.float32), self.pose_dtype)  # Keras casts to fp32
            for k, i in enumerate(idx):
                self.tracks[i].keypoints = np.asarray(kp[k], dtype=np.float64)

    @staticmethod
    def _snapshot(tr: Track) -> dict:
        return {"id": tr.id, "x": tr.x.copy(), "P": tr.P.copy(), "lifetime": float(tr.lifetime),
                "spread_est": tr.spread_est.copy(), "group_disp_est": tr.group_disp_est.copy(),
                "N_est": float(tr.N_est), "point_num": int(tr.point_num), "centroid": tr.centroid.copy(),
                "min_vals": tr.min_vals.copy(), "max_vals": tr.max_vals.copy(), "static": bool(tr.static),
                "ring_counts": np.array([len(f) for f in tr.ring.frames], dtype=np.int32),
                "keypoints": np.asarray(tr.keypoints, dtype=np.float64).copy()}
