"""Configuration module with the names the reference's ``constants.py`` exposes for the hot path.

Edit the values here (or assign to the module attributes at run time) exactly as the reference's README asks
its users to do; ``Tracking.TrackBuffer`` snapshots them into an ``mmw_config`` when it creates its device
context.  Only what the per-frame perception path reads is present -- GUI, logging and serial-port settings of
the reference are out of scope.  Values are the reference defaults (checked against the reference's own module
by tests/test_abi_cpu.py through tests/golden/known_answers.json).
"""
import numpy as np

# smart window: sensitive object and fade-square sizing (constants.py:31-33, 50-54)
M_X = 0.32
M_Y = -0.6
M_Z = 1.3
V_SCREEN_FADE_SIZE_MAX = 0.3
V_SCREEN_FADE_SIZE_MIN = 0.2
V_SCREEN_FADE_WEIGHT = 0.08

# Doppler resolution of the radar profile [m/s per index] (dopplerResolutionMps, ReadDataIWR1443.py:249-257).  Not a
# constant of the reference (it derives it from the .cfg it uploads); None = take it from the data (Utils._doppler_units)
DOPPLER_RESOLUTION = None

# sensor pose (constants.py:41-42)
S_HEIGHT = 1.8
S_TILT = -5

# replay / logging (constants.py:58-61)
FB_EXPERIMENT_FILE_SIZE = 200
FB_READ_BUFFER_SIZE = 40

# clustering (constants.py:66-73)
FB_FRAMES_BATCH = 2
FB_FRAMES_BATCH_STATIC = 2
DB_Z_WEIGHT = 0.4
DB_RANGE_WEIGHT = 0.03
DB_EPS = 0.3
DB_MIN_SAMPLES_MIN = 35

# tracking and Kalman (constants.py:85-104)
TR_MAX_TRACKS = 4
TR_LIFETIME_DYNAMIC = 3
TR_LIFETIME_STATIC = 7
TR_VEL_THRES = 0.12
TR_GATE = 4.5
KF_R_STD = 0.1
KF_Q_STD = 1
KF_P_INIT = 0.1
KF_GROUP_DISP_EST_INIT = 0.1
KF_ENABLE_EST = False
KF_A_N = 0.9
KF_EST_POINTNUM = 10
KF_SPREAD_LIM = [0.2, 0.2, 2, 1.2, 1.2, 0.2]
KF_A_SPR = 0.9

# pose model (constants.py:108-172)
INTENSITY_MU = 27.0187
INTENSITY_STD = 70.351
MODEL_MIN_INPUT = 0
P_MODEL_PATH = "../trained_cases/Our_system/model/MARS.h5"
MODEL_DEFAULT_POSTURE = np.array(
    "0.0000 -0.0007 -0.0006 -0.0038 -0.1820 -0.2540 -0.2579 0.1830 0.2957 0.2940 -0.0805 -0.1141 -0.1232 -0.1358 "
    "0.0796 0.1436 0.1558 0.1720 -0.0007 0.7699 1.0906 1.4020 1.5513 1.2893 1.0360 0.7994 1.2865 1.0483 0.8117 "
    "0.7670 0.3428 0.0000 -0.0746 0.7713 0.3706 -0.0128 -0.0796 1.3255 0.0752 0.0533 0.0203 0.0000 0.0496 0.1350 "
    "0.1303 0.0345 0.1277 0.1050 0.0392 0.0533 0.0786 -0.0056 0.0346 -0.0007 0.0683 -0.0082 0.0312".split(),
    dtype=np.float64)


class CONST_ACC_MODEL:
    """Constant-acceleration motion model, state [x y z vx vy vz ax ay az] (constants.py:176-215).  The matrices
    are provided for callers that inspect them; the device kernels build F(dt) and Q(dt) themselves."""
    KF_DIM = [9, 6]
    KF_H = np.hstack([np.eye(6), np.zeros((6, 3))])

    @staticmethod
    def STATE_VEC(init):
        return [*init[:6], 0, 0, 0]

    @staticmethod
    def KF_F(dt):
        F = np.eye(9)
        F[np.arange(6), np.arange(6) + 3] = dt
        F[np.arange(3), np.arange(3) + 6] = 0.5 * dt ** 2
        return F

    @staticmethod
    def KF_Q_DISCR(dt):
        qw = np.array([[dt ** 4 / 4, dt ** 3 / 2, dt ** 2 / 2], [dt ** 3 / 2, dt ** 2, dt], [dt ** 2 / 2, dt, 1.0]])
        return np.kron(np.eye(3), qw) * KF_Q_STD


MOTION_MODEL = CONST_ACC_MODEL
