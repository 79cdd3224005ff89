"""Numbers of the ReVolt DP hot path, mirrored from ml4ca_b200/csrc/ml4ca_constants.h.

TEST INFRASTRUCTURE (see oracle/__init__.py).  tests/test_constants.py parses the C header and
checks every entry below against it.  Reference provenance is listed in the header.
"""
import math

PI = math.pi

# thrusters, allocator order [port, star, bow]   (qp_allocator.py:51-55,69-70)
LX = (-1.12, -1.12, 1.08)
LY = (-0.15, 0.15, 0.0)
K_THRUST = (0.00205, 0.00205, 0.0009)
F_MAX = (20.5, 20.5, 9.0)
BOW_ANGLE_FIXED = PI / 2.0

# SLSQP allocator (qp_allocator.py:57-58,196-200,232,307)
QP_DF = (5.0, 5.0, 2.0)
QP_DA = (PI / 12.0, PI / 12.0)
QP_ALPHA_BOUND = 2.0 * PI
QP_SLACK_BOUND = 1.0
QP_W_RATE = 0.25
QP_CLEAN_EPS = 0.01
BOW_THROTTLE_GAIN = 2.5

# stand-in hull (DECLARED; absent from the reference)
M11, M22, M33 = 264.0, 306.0, 322.0
XU, XUU = 10.0, 13.8
YV, YVV = 100.0, 222.0
NR, NRR = 60.0, 90.3
SIM_DT = 0.01
# hull_model -> (m11, m22, m33, Xu, Xuu, Yv, Yvv, Nr, Nrr); model 1 = tools/sysid_hull.py --constrained --wrench-lag
HULL_MODELS = {
    0: dict(m11=M11, m22=M22, m33=M33, Xu=XU, Xuu=XUU, Yv=YV, Yvv=YVV, Nr=NR, Nrr=NRR),
    1: dict(m11=271.0, m22=316.0, m33=320.0, Xu=12.3, Xuu=12.13, Yv=0.0, Yvv=555.6, Nr=106.0, Nrr=1.78),
}
H1_LAG_S = 0.92

# env (customEnv.py:26,79-83,386-399)
N_SUBSTEPS = 20
MAX_EP_LEN = 400
SS_BOUNDS = (8.0, 8.0, 45.0 * PI / 180.0, 1.4, 0.30, 0.52)
THRUST_BOUND = 100.0
VEL_FRACTION = 0.30

# reward (customEnv.py:78,86-88,263)
REW_VEL_C = (0.5, 0.5, 1.0)
REW_SIGMA_POS = 1.0
REW_SIGMA_YAW = 5.0
REW_THRUST_C = (0.20, 0.30, 0.30)       # env order [bow, port, star]
REW_DTHRUST_C = (0.05, 0.05, 0.05)
REW_DANGLE_C = (0.0, 0.01, 0.01)        # env order [bow, port, star]

# pseudoinverse + PID baseline (gains DECLARED by this build; saturation from SupervisedTau.py:37)
PID_KP = (30.0, 30.0, 60.0)
PID_KD = (90.0, 120.0, 120.0)
PID_KI = (1.0, 1.0, 2.0)
PID_SAT = (69.0, 30.0, 80.0)
PID_DT = 0.2
