"""Float64 NumPy integration of the DECLARED stand-in 3-DOF hull, plus a DigiTwin duck type.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference steps the proprietary Java "Cybersea" simulator through py4j
(specific/digitwin.py:213-217, customEnv.py:124); the simulator and every hull parameter are
absent from the repository.  The build therefore states its own equations -- THIS GAP IS DECLARED,
PARITY OF THE DYNAMICS IS UNPINNED -- and checks the CUDA integrator against this float64
integration of the same equations:

    tau   = sum_i  F_i [cos a_i, sin a_i, lx_i sin a_i - ly_i cos a_i],   F_i = K_i n_i |n_i|
            (B(alpha): src/qp/qp.py:24-36, SupervisedTau.py:42-52;  K: qp_allocator.py:51-55)
    m11 du/dt = tau_X + m22 v r - (Xu + Xuu|u|) u
    m22 dv/dt = tau_Y - m11 u r - (Yv + Yvv|v|) v
    m33 dr/dt = tau_N - (m22 - m11) u v - (Nr + Nrr|r|) r
    dN/dt = cos(psi) u - sin(psi) v ;  dE/dt = sin(psi) u + cos(psi) v ;  dpsi/dt = r

integrated with the semi-implicit Euler scheme (h = 10 ms, customEnv.py:79-81)

    nu+ = nu + h nu_dot(nu) ;  N+,E+ = N,E + h R(psi) nu+ ;  psi+ = psi + h r+

Commanded thrust/azimuth act instantly and stay constant over the sub-steps of one env step -- unless an
actuator lag T > 0 is declared (ml4ca_env_cfg.actuator_lag_s): the wrench then follows its command through

    tau_act+ = tau_act + h / (T + h) (tau_cmd - tau_act)          (implicit Euler of  T d tau_act/dt = tau_cmd - tau_act)

advanced at the start of every sub-step, and tau_act is a state of the env (zero after a reset).
``params`` selects the DECLARED parameter set (constants.HULL_MODELS; 0 = default).
"""
import numpy as np

from . import constants as C

# env thruster order [bow, port, star]  ->  allocator order [port, star, bow] (customEnv.py:47-53)
_LX_ENV = (C.LX[2], C.LX[0], C.LX[1])
_LY_ENV = (C.LY[2], C.LY[0], C.LY[1])
_K_ENV = (C.K_THRUST[2], C.K_THRUST[0], C.K_THRUST[1])


def thruster_wrench(n_pct, alpha):
    """tau [3, ...] from thrust commands n_pct [3, ...] (%) and azimuths alpha [3, ...] in env order."""
    n_pct = np.asarray(n_pct, dtype=np.float64)
    alpha = np.asarray(alpha, dtype=np.float64)
    tx = 0.0
    ty = 0.0
    tn = 0.0
    for i in range(3):
        f = _K_ENV[i] * n_pct[i] * np.abs(n_pct[i])
        c, s = np.cos(alpha[i]), np.sin(alpha[i])
        tx = tx + f * c
        ty = ty + f * s
        tn = tn + f * (_LX_ENV[i] * s - _LY_ENV[i] * c)
    return np.stack([tx, ty, tn])


def integrate(eta, nu, tau, n_sub, h=C.SIM_DT, params=None, tau_act=None, lag_s=0.0):
    """n_sub semi-implicit Euler sub-steps.  eta, nu, tau: float64 [3, ...].  Returns (eta, nu), or (eta, nu, tau_act)
    when a lagged wrench state ``tau_act`` [3, ...] is passed."""
    p = C.HULL_MODELS[0] if params is None else params
    m11, m22, m33 = p['m11'], p['m22'], p['m33']
    N, E, psi = (np.array(x, dtype=np.float64) for x in eta)
    u, v, r = (np.array(x, dtype=np.float64) for x in nu)
    cmd = np.stack([np.asarray(x, dtype=np.float64) for x in tau])
    lagged = tau_act is not None
    act = np.array(tau_act, dtype=np.float64) if lagged else cmd
    k = h / (lag_s + h) if lag_s > 0 else 1.0
    for _ in range(int(n_sub)):
        if lagged:
            act = act + k * (cmd - act)
        tx, ty, tn = act
        du = (tx + m22 * v * r - (p['Xu'] + p['Xuu'] * np.abs(u)) * u) / m11
        dv = (ty - m11 * u * r - (p['Yv'] + p['Yvv'] * np.abs(v)) * v) / m22
        dr = (tn - (m22 - m11) * u * v - (p['Nr'] + p['Nrr'] * np.abs(r)) * r) / m33
        u = u + h * du
        v = v + h * dv
        r = r + h * dr
        c, s = np.cos(psi), np.sin(psi)
        N = N + h * (c * u - s * v)
        E = E + h * (s * u + c * v)
        psi = psi + h * r
    if lagged:
        return np.stack([N, E, psi]), np.stack([u, v, r]), act
    return np.stack([N, E, psi]), np.stack([u, v, r])


class VesselTwin(object):
    """Duck type of ``DigiTwin`` (digitwin.py:50,213): ``val(module, feat, val=None)`` and ``step(n)``.

    Lets the UNMODIFIED reference env (customEnv.py) run on the stand-in hull.  ``frozen=True``
    makes ``step`` a no-op (a null simulator), which isolates the wrapper arithmetic.
    Features used by the reference: Hull.{Eta,Nu,Yaw,PosNED,PosAttitude,VelocityNu,StateResetOn},
    THR{1,2,3}.{ThrustOrTorqueCmdMtc,AzmCmdMtc,MtcOn}, THR1.LinActuator.
    While Hull.StateResetOn == 1 the hull state is held (reset's 50 settle steps, customEnv.py:164-167).
    """

    def __init__(self, frozen=False):
        self.frozen = frozen
        self.eta = np.zeros(3)
        self.nu = np.zeros(3)
        self.thrust = np.zeros(3)            # env order [bow, port, star], percent
        self.azimuth = np.array([C.BOW_ANGLE_FIXED, 0.0, 0.0])
        self.reset_on = 0
        self.misc = {}

    def val(self, module, feat, val=None, report=False):
        if module == 'Hull':
            if val is None:
                if feat == 'Eta':
                    return [self.eta[0], self.eta[1], 0.0, 0.0, 0.0, self.eta[2]]
                if feat == 'Nu':
                    return [self.nu[0], self.nu[1], 0.0, 0.0, 0.0, self.nu[2]]
                if feat == 'Yaw':
                    return self.eta[2]
                return self.misc.get((module, feat), 0.0)
            if feat == 'PosNED':
                self.eta[0], self.eta[1] = float(val[0]), float(val[1])
            elif feat == 'PosAttitude':
                self.eta[2] = float(val[2])
            elif feat == 'VelocityNu':
                self.nu[:] = [float(val[0]), float(val[1]), float(val[5])]
            elif feat == 'StateResetOn':
                self.reset_on = int(val)
            else:
                self.misc[(module, feat)] = val
            return None
        if module in ('THR1', 'THR2', 'THR3'):
            i = int(module[3]) - 1
            if feat == 'ThrustOrTorqueCmdMtc':
                if val is None:
                    return self.thrust[i]
                self.thrust[i] = float(val)
            elif feat == 'AzmCmdMtc':
                if val is None:
                    return self.azimuth[i]
                self.azimuth[i] = float(val)
            else:
                if val is None:
                    return self.misc.get((module, feat), 0.0)
                self.misc[(module, feat)] = val
            return None
        if val is None:
            return self.misc.get((module, feat), 0.0)
        self.misc[(module, feat)] = val
        return None

    def step(self, steps=1):
        if self.frozen or self.reset_on:
            return
        tau = thruster_wrench(self.thrust, self.azimuth)
        self.eta, self.nu = integrate(self.eta, self.nu, tau, steps)
