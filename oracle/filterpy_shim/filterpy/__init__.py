"""TEST INFRASTRUCTURE ONLY -- restatement of the filterpy 1.4.5 API surface the
reference uses (requirements.txt:1 pins filterpy==1.4.5; it is not installed in
this image and there is no network).

Only what Tracking.py:5,74-97,381-384,393 and constants.py:2,212-214 touch is
restated, from filterpy's published semantics:
``filterpy.kalman.KalmanFilter`` (__init__/predict/update) and
``filterpy.common.Q_discrete_white_noise``.  It exists so that the reference's
own, unmodified modules import and run when oracle/ref_harness.py generates
golden vectors.  Nothing in the product imports this package.
"""
__version__ = "1.4.5-shim"
