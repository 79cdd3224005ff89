"""ml4ca_constants.h (kernels) and oracle/constants.py (checker) must carry the same numbers."""
import math
import os
import re

from oracle import constants as C

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ml4ca_b200", "csrc", "ml4ca_constants.h")


def parse_header():
    env = {}
    pat = re.compile(r"^#define\s+(ML4CA_\w+)\s+(.+?)\s*(?://.*)?$")
    for line in open(HEADER):
        m = pat.match(line.strip())
        if not m or m.group(1).endswith("_H_"):
            continue
        expr = re.sub(r"ML4CA_(\w+)", lambda k: "env['ML4CA_%s']" % k.group(1), m.group(2))
        env[m.group(1)] = eval(expr, {"env": env})
    return env


def test_header_matches_oracle():
    h = parse_header()
    assert h["ML4CA_PI"] == math.pi
    assert (h["ML4CA_LX_PORT"], h["ML4CA_LX_STAR"], h["ML4CA_LX_BOW"]) == C.LX
    assert (h["ML4CA_LY_PORT"], h["ML4CA_LY_STAR"], h["ML4CA_LY_BOW"]) == C.LY
    assert (h["ML4CA_K_STERN"], h["ML4CA_K_STERN"], h["ML4CA_K_BOW"]) == C.K_THRUST
    assert (h["ML4CA_FMAX_STERN"], h["ML4CA_FMAX_STERN"], h["ML4CA_FMAX_BOW"]) == C.F_MAX
    assert h["ML4CA_BOW_ANGLE_FIXED"] == C.BOW_ANGLE_FIXED
    assert (h["ML4CA_QP_DF_STERN"], h["ML4CA_QP_DF_STERN"], h["ML4CA_QP_DF_BOW"]) == C.QP_DF
    assert (h["ML4CA_QP_DA_STERN"],) * 2 == C.QP_DA
    assert h["ML4CA_QP_ALPHA_BOUND"] == C.QP_ALPHA_BOUND and h["ML4CA_QP_SLACK_BOUND"] == C.QP_SLACK_BOUND
    assert h["ML4CA_QP_W_RATE"] == C.QP_W_RATE and h["ML4CA_QP_CLEAN_EPS"] == C.QP_CLEAN_EPS
    assert h["ML4CA_BOW_THROTTLE_GAIN"] == C.BOW_THROTTLE_GAIN
    assert (h["ML4CA_M11"], h["ML4CA_M22"], h["ML4CA_M33"]) == (C.M11, C.M22, C.M33)
    assert (h["ML4CA_XU"], h["ML4CA_XUU"], h["ML4CA_YV"], h["ML4CA_YVV"], h["ML4CA_NR"], h["ML4CA_NRR"]) == \
        (C.XU, C.XUU, C.YV, C.YVV, C.NR, C.NRR)
    h1 = C.HULL_MODELS[1]
    assert [h["ML4CA_H1_" + k.upper()] for k in ("m11", "m22", "m33", "Xu", "Xuu", "Yv", "Yvv", "Nr", "Nrr")] == \
        [h1[k] for k in ("m11", "m22", "m33", "Xu", "Xuu", "Yv", "Yvv", "Nr", "Nrr")]
    assert h["ML4CA_H1_LAG_S"] == C.H1_LAG_S
    assert h["ML4CA_SIM_DT"] == C.SIM_DT and h["ML4CA_N_SUBSTEPS"] == C.N_SUBSTEPS
    assert h["ML4CA_MAX_EP_LEN"] == C.MAX_EP_LEN
    assert (h["ML4CA_BOUND_POS"], h["ML4CA_BOUND_POS"], h["ML4CA_BOUND_YAW"], h["ML4CA_BOUND_U"], h["ML4CA_BOUND_V"],
            h["ML4CA_BOUND_R"]) == C.SS_BOUNDS
    assert h["ML4CA_THRUST_BOUND"] == C.THRUST_BOUND and h["ML4CA_VEL_FRACTION"] == C.VEL_FRACTION
    assert (h["ML4CA_REW_VEL_CU"], h["ML4CA_REW_VEL_CV"], h["ML4CA_REW_VEL_CR"]) == C.REW_VEL_C
    assert (h["ML4CA_REW_SIGMA_POS"], h["ML4CA_REW_SIGMA_YAW"]) == (C.REW_SIGMA_POS, C.REW_SIGMA_YAW)
    assert (h["ML4CA_REW_THRUST_C_BOW"], h["ML4CA_REW_THRUST_C_STERN"], h["ML4CA_REW_THRUST_C_STERN"]) == C.REW_THRUST_C
    assert (h["ML4CA_REW_DTHRUST_C"],) * 3 == C.REW_DTHRUST_C
    assert (h["ML4CA_REW_DANGLE_C_BOW"], h["ML4CA_REW_DANGLE_C_STERN"], h["ML4CA_REW_DANGLE_C_STERN"]) == C.REW_DANGLE_C
    assert (h["ML4CA_PID_KP_X"], h["ML4CA_PID_KP_Y"], h["ML4CA_PID_KP_N"]) == C.PID_KP
    assert (h["ML4CA_PID_KD_X"], h["ML4CA_PID_KD_Y"], h["ML4CA_PID_KD_N"]) == C.PID_KD
    assert (h["ML4CA_PID_KI_X"], h["ML4CA_PID_KI_Y"], h["ML4CA_PID_KI_N"]) == C.PID_KI
    assert (h["ML4CA_PID_SAT_X"], h["ML4CA_PID_SAT_Y"], h["ML4CA_PID_SAT_N"]) == C.PID_SAT
    assert h["ML4CA_PID_DT"] == C.PID_DT


def test_hull_calibration_matches_reference_top_speeds():
    """customEnv.py:13-18 ('with thrust losses', full): +1.4 m/s, 0.30 m/s, 0.52 rad/s at full thrust."""
    fx = 2 * C.F_MAX[0]
    fy = 2 * C.F_MAX[0] + C.F_MAX[2]
    mz = 2 * C.F_MAX[0] * abs(C.LX[0]) + C.F_MAX[2] * C.LX[2]
    for model, p in C.HULL_MODELS.items():          # both declared parameter sets keep the stated top speeds
        assert abs(p["Xu"] * 1.4 + p["Xuu"] * 1.4 ** 2 - fx) < 0.1, model
        assert abs(p["Yv"] * 0.30 + p["Yvv"] * 0.30 ** 2 - fy) < 0.1, model
        assert abs(p["Nr"] * 0.52 + p["Nrr"] * 0.52 ** 2 - mz) < 0.2, model
