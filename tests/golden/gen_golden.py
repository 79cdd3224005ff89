"""Generate the golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE CODE.

Only works in the build container (needs the read-only checkout at /root/reference); the resulting
.npz files are committed so that the GPU box -- where the reference does not exist -- can check the
oracle and the CUDA path against the reference's own outputs.

    python tests/golden/gen_golden.py

env_<kind>_<mode>.npz : reference customEnv classes driven through the DigiTwin duck type
                        oracle.vessel.VesselTwin (float64).  mode 'null' = frozen simulator (wrapper
                        arithmetic only), 'hull' = the declared stand-in hull integrated in float64.
qp_config1.npz        : reference QPTA.solve_QP + tau_controller_callback_func post-processing on the
                        SURVEY.md section 8(d) config-1 batch at its stated size (4096 demands; container SciPy, see
                        oracle/qp_oracle.py header), rospy.get_time() pinned to 0; plus, per demand, SLSQP's exit
                        status / iteration count, the raw x and the converged KKT point of the reference's basin
                        (oracle.qp_oracle.solve_converged) with its objective and active set; the same call with tau
                        moved by 1e-9 (*_pert) and with analytic instead of finite-difference derivatives (*_exact).
                        Inputs are rounded to fp32-representable values (the CUDA path takes fp32 rows).
qp_switches.npz       : reference QPTA.solve_QP with weight_matrix / reduce_fuel / reduce_flickering / reduce_angular
                        (qp_allocator.py:108,116-150), 96 demands per case.
policy_*.npz          : the shipped TF1 checkpoints' actor/critic weights, read by ml4ca_b200/tf_checkpoint.py
ros_adapter.npz       : the deployment node RLTA (src/rl/ROS/rl_allocator/src/rl_allocator.py) driven message by message
                        with a stub actor: state vector, ROS-order action, published message fields.
resetacts_final.npz   : RevoltFinal(reset_acts=True): reset observation with the N(0, 0.1) previous thrust
                        (customEnv.py:179-188) and the first steps after it.
error_frame.npz       : the reference's ErrorFrame class on poses with |heading| and |heading error| up to 400 (the deg=True
                        quirk of wrap_angle on radians only wraps beyond +-180).
gae.npz               : the reference's own TrajectoryBuffer (ppo.py:21-105) with core.discount_cumsum and
                        mpi_statistics_scalar, imported unmodified behind tensorflow / gym / mpi4py stubs: store(),
                        finish_path(last_val) per path (died / cut / epoch end), get().
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_loader, vessel, qp_oracle  # noqa: E402
import scipy  # noqa: E402
import scipy.signal  # noqa: E402


def gen_env(kind, cls_name, kwargs, frozen, B, T, seed):
    mod = ref_loader.load_env_module()
    rng = np.random.default_rng(seed)
    act_dim = None
    rec = {k: [] for k in ('eta0', 'nu0', 'actions', 'obs0', 'obs', 'rew', 'done', 'eta', 'nu')}
    for b in range(B):
        twin = vessel.VesselTwin(frozen=frozen)
        if cls_name == 'Revolt':
            env = mod.Revolt(digitwin=twin, real_ss_bounds=[8.0, 8.0, np.pi / 2, 1.4, 0.30, 0.52], **kwargs)
        else:
            env = getattr(mod, cls_name)(twin, **kwargs)
        act_dim = env.num_actions
        eta0 = rng.uniform(-1, 1, 3) * np.array([7.0, 7.0, 0.7])
        nu0 = rng.uniform(-1, 1, 3) * np.array([0.5, 0.12, 0.2])
        init = {'Hull.PosNED': [eta0[0], eta0[1]], 'Hull.PosAttitude': [0, 0, eta0[2]],
                'Hull.VelocityNu': [nu0[0], nu0[1], 0, 0, 0, nu0[2]]}
        o0 = env.reset(**init)
        acts, obs, rew, done, etas, nus = [], [], [], [], [], []
        for t in range(T):
            a = rng.uniform(-1.25, 1.25, act_dim)
            if t % 7 == 3:
                a[rng.integers(0, act_dim)] = 1.0      # exactly on the clip boundary
            o, r, d, _ = env.step(a)
            acts.append(a); obs.append(np.array(o, dtype=np.float64))
            rew.append(float(np.asarray(r).ravel()[0])); done.append(bool(d))
            etas.append(twin.eta.copy()); nus.append(twin.nu.copy())
        rec['eta0'].append(eta0); rec['nu0'].append(nu0); rec['obs0'].append(np.array(o0, dtype=np.float64))
        rec['actions'].append(acts); rec['obs'].append(obs); rec['rew'].append(rew); rec['done'].append(done)
        rec['eta'].append(etas); rec['nu'].append(nus)
        max_ep_len, dt = env.max_ep_len, env.dt
    out = {
        'eta0': np.array(rec['eta0']).T,                               # [3, B]
        'nu0': np.array(rec['nu0']).T,
        'obs0': np.array(rec['obs0']).T,                               # [obs, B]
        'actions': np.transpose(np.array(rec['actions']), (1, 2, 0)),  # [T, act, B]
        'obs': np.transpose(np.array(rec['obs']), (1, 2, 0)),          # [T, obs, B]
        'rew': np.array(rec['rew']).T,                                 # [T, B]
        'done': np.array(rec['done']).T,
        'eta': np.transpose(np.array(rec['eta']), (1, 2, 0)),          # [T, 3, B]
        'nu': np.transpose(np.array(rec['nu']), (1, 2, 0)),
        'max_ep_len': np.array(max_ep_len), 'dt': np.array(dt),
    }
    return out


def _qp_chunk(args):
    """One worker: the UNMODIFIED reference QPTA on demands [lo, hi) + the converged point of its basin."""
    lo, hi, n, seed = args
    qp = ref_loader.load_qp_module()
    ta = qp.QPTA()
    tau, prev = _qp_inputs(n, seed)
    rows = []
    for j in range(lo, hi):
        ta.previous_thruster_state = [float(v) for v in prev[:, j]] + [np.pi / 2]
        x, ok = ta.solve_QP(tau[:, j].reshape(3, 1))
        # full callback (post-processing :267-320) on a fresh object state
        ta.previous_thruster_state = [float(v) for v in prev[:, j]] + [np.pi / 2]
        ta.tau_controller_callback_func(ref_loader.wrench(tau[0, j], tau[1, j], tau[2, j]))
        eff = [ta.pub_stern_thruster_setpoints.last.port_effort, ta.pub_stern_thruster_setpoints.last.star_effort]
        ang = [ta.pub_stern_angles.last.port, ta.pub_stern_angles.last.star]
        bow = float(ta.pub_bow_control.last.throttle_bow)
        newprev = list(ta.previous_thruster_state)
        # the restated stock solve (bit-identical, tests/test_oracle_pinning.py) exposes what solve_QP hides: the raw x,
        # SLSQP's exit status and iteration count
        xs, oks, raw, info = qp_oracle.solve_stock(tau[:, j], prev[:, j], return_info=True)
        assert oks == bool(ok) and np.array_equal(xs, np.array(x))
        # converged KKT point of the basin the reference lands in (float64; what the 1e-5 tolerance is measured against)
        xc, okc, rc = qp_oracle.solve_converged(tau[:, j], prev[:, j])
        viol_c = qp_oracle.kkt_residual(xc, tau[:, j], prev[:, j])[1] if okc else np.inf
        # how sharply the reference itself depends on its input: the same call with tau moved by 1e-9 (relative), far below
        # the fp32 resolution of the demand.  Rows where that changes the flag or the end point are rows no independent
        # implementation can be expected to reproduce.
        _, okp, rawp = qp_oracle.solve_stock(tau[:, j] * (1.0 + 1e-9), prev[:, j])
        # ... and how much it owes to its finite-difference derivatives: the same call with analytic derivatives
        xa, oka, nita = qp_oracle.solve_stock_exact_derivatives(tau[:, j], prev[:, j])
        rows.append((np.array(x), bool(ok), eff, ang, bow, newprev, raw, info['status'], info['nit'],
                     xc, rc, viol_c, qp_oracle.objective(raw, prev[:, j]), qp_oracle.objective(xc, prev[:, j]),
                     qp_oracle.active_set(xc, prev[:, j], tol=2e-5), qp_oracle.active_set(raw, prev[:, j], tol=1e-5),
                     bool(okp), rawp, xa, bool(oka), nita))
    return lo, rows


QP_SWITCH_CASES = {   # solve_QP(tau_d, weight_matrix, reduce_fuel, reduce_flickering, reduce_angular), :108
    'nofuel': dict(reduce_fuel=False),
    'noflick': dict(reduce_flickering=False),
    'noang': dict(reduce_angular=False),
    'bare': dict(reduce_fuel=False, reduce_flickering=False, reduce_angular=False),
    'weighted': dict(weight_matrix=np.diag([2.0, 1.0, 0.5, 1.0, 0.5, 2.0, 0.1, 0.4, 0.3, 0.2, 0.5])),
}


def gen_qp_switches(n=96, seed=9):
    """The reference's QPTA.solve_QP with its objective switches, per case on n demands of the config-1 law (fp32-
    representable inputs): cleaned x, success, and the restated stock / exact-derivative solves for the raw x."""
    qp = ref_loader.load_qp_module()
    ta = qp.QPTA()
    tau, prev = _qp_inputs(n, seed)
    out = {'tau': tau, 'prev': prev}
    for tag, kw in QP_SWITCH_CASES.items():
        w, fuel = qp_oracle.weights_from_switches(**kw)
        xs, oks, raws, xes, okes = [], [], [], [], []
        for j in range(n):
            ta.previous_thruster_state = [float(v) for v in prev[:, j]] + [np.pi / 2]
            x, ok = ta.solve_QP(tau[:, j].reshape(3, 1), **kw)
            xo, oko, raw = qp_oracle.solve_stock(tau[:, j], prev[:, j], w=w, fuel=fuel)
            assert oko == bool(ok) and np.array_equal(xo, np.array(x)), (tag, j)      # the restatement IS the reference
            xe, oke, _ = qp_oracle.solve_stock_exact_derivatives(tau[:, j], prev[:, j], w=w, fuel=fuel)
            xs.append(np.array(x)); oks.append(bool(ok)); raws.append(raw); xes.append(xe); okes.append(oke)
        out[tag + '__x'] = np.array(xs).T
        out[tag + '__success'] = np.array(oks)
        out[tag + '__x_raw'] = np.array(raws).T
        out[tag + '__x_raw_exact'] = np.array(xes).T
        out[tag + '__success_exact'] = np.array(okes)
        out[tag + '__weights'] = w
        out[tag + '__fuel'] = np.array(fuel)
    return out


def _qp_inputs(n, seed):
    """SURVEY 8(d) config 1, rounded to fp32-representable values: the CUDA path takes fp32 rows, and the reference is
    run on exactly those numbers."""
    tau, prev = qp_oracle.synth_batch(n, seed=seed)
    return tau.astype(np.float32).astype(np.float64), prev.astype(np.float32).astype(np.float64)


def gen_qp(n, seed, workers=None):
    import multiprocessing as mp
    tau, prev = _qp_inputs(n, seed)
    workers = workers or os.cpu_count() or 1
    step = max(1, (n + 4 * workers - 1) // (4 * workers))
    jobs = [(lo, min(lo + step, n), n, seed) for lo in range(0, n, step)]
    with mp.get_context("fork").Pool(workers) as pool:
        parts = sorted(pool.map(_qp_chunk, jobs), key=lambda t: t[0])
    rows = [r for _, part in parts for r in part]
    col = lambda i: [r[i] for r in rows]   # noqa: E731
    return {'tau': tau, 'prev': prev, 'x': np.array(col(0)).T, 'success': np.array(col(1)),
            'stern_effort': np.array(col(2)).T, 'pod_angle_deg': np.array(col(3)).T, 'bow_throttle': np.array(col(4)),
            'new_prev': np.array(col(5)).T, 'x_raw': np.array(col(6)).T,
            'slsqp_status': np.array(col(7), dtype=np.int32), 'slsqp_nit': np.array(col(8), dtype=np.int32),
            'x_conv': np.array(col(9)).T, 'kkt_conv': np.array(col(10)), 'viol_conv': np.array(col(11)),
            'obj_raw': np.array(col(12)), 'obj_conv': np.array(col(13)),
            'mask_conv': np.array(col(14), dtype=np.int32), 'mask_raw': np.array(col(15), dtype=np.int32),
            'success_pert': np.array(col(16)), 'x_raw_pert': np.array(col(17)).T,
            'x_raw_exact': np.array(col(18)).T, 'success_exact': np.array(col(19)),
            'nit_exact': np.array(col(20), dtype=np.int32),
            'scipy_version': np.array(scipy.__version__)}


def gen_gae(seed):
    """The reference's OWN TrajectoryBuffer (ppo.py:21-105; core.discount_cumsum core.py:48-63; mpi_statistics_scalar
    mpi_tools.py:71-93), imported unmodified behind tensorflow / gym / mpi4py stubs (oracle.ref_loader.load_ppo_module):
    one buffer of 64 steps filled through store(), closed path by path through finish_path(last_val) -- a died episode
    (last_val 0), two cut ones (last_val = V) and the epoch end -- then get()."""
    ppo = ref_loader.load_ppo_module()
    rng = np.random.default_rng(seed)
    T = 64
    gamma, lam = 0.99, 0.97
    buf = ppo.TrajectoryBuffer(9, 7, T, gamma, lam)
    rews = rng.normal(size=T).astype(np.float32)
    vals = rng.normal(size=T).astype(np.float32)
    ends = {17: 0.0, 30: 0.41, 46: -0.83, 63: 0.37}       # path end (inclusive step) -> last_val handed to finish_path
    flags = np.zeros(T, dtype=np.uint8)                     # this build's flag byte: 1 terminal, 2 episode-length cut
    boot = np.zeros(T, dtype=np.float32)
    for t in range(T):
        buf.store(rng.normal(size=9), rng.normal(size=7), rews[t], vals[t], 0.0)
        if t in ends:
            buf.finish_path(ends[t])
            if t != T - 1:
                flags[t] = 1 if ends[t] == 0.0 else 2
                boot[t] = ends[t]
    adv_raw, ret = buf.adv_buf.copy(), buf.ret_buf.copy()
    obs, act, adv_norm, ret2, logp = buf.get()
    # the single-path case of round 1 (one finish_path at the end)
    buf1 = ppo.TrajectoryBuffer(9, 7, 37, gamma, lam)
    r1, v1 = rng.normal(size=37).astype(np.float32), rng.normal(size=37).astype(np.float32)
    for t in range(37):
        buf1.store(np.zeros(9), np.zeros(7), r1[t], v1[t], 0.0)
    buf1.finish_path(np.float32(0.37))
    return {'rews': r1, 'vals': v1, 'last_val': np.float32(0.37), 'gamma': np.array(gamma), 'lam': np.array(lam),
            'adv': buf1.adv_buf.copy(), 'ret': buf1.ret_buf.copy(),
            'multi_rews': rews, 'multi_vals': vals, 'multi_flags': flags, 'multi_boot': boot,
            'multi_last_val': np.float32(ends[T - 1]), 'multi_adv': adv_raw, 'multi_ret': ret,
            'multi_adv_normalized': np.asarray(adv_norm, dtype=np.float32)}


def gen_error_frame(n=512, seed=41):
    """The reference's ErrorFrame itself (errorFrame.py:4-38; rotation_matrix / wrap_angle mathematics.py:7-17) on poses
    whose heading and heading error run far beyond +-pi and beyond +-180 (where the deg=True quirk of wrap_angle, applied to
    radians, finally wraps)."""
    ef_mod = ref_loader.load_error_frame_module()
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-1, 1, (3, n)) * np.array([[8.0], [8.0], [4.0]])
    ref = rng.uniform(-1, 1, (3, n)) * np.array([[8.0], [8.0], [4.0]])
    big = slice(n // 2, None)
    pos[2, big] = rng.uniform(-400.0, 400.0, n - n // 2)            # |psi| >= 180 included
    ref[2, big] = rng.uniform(-400.0, 400.0, n - n // 2)
    pos[2, n // 2] = 180.0                                           # exactly on the wrap point: (180 + 180) mod 360 - 180
    ref[2, n // 2] = 0.0
    pos, ref = pos.astype(np.float32).astype(np.float64), ref.astype(np.float32).astype(np.float64)
    err = np.zeros((3, n))
    for j in range(n):
        ef = ef_mod.ErrorFrame(pos=list(pos[:, j]), ref=list(ref[:, j]))
        err[:, j] = np.asarray(ef.get_pose(), dtype=np.float64).ravel()
    return {'pos': pos, 'ref': ref, 'err': err}


def gen_policy():
    """Shipped actor/critic weights (TF1 bundle, read without TensorFlow) as flat parameter vectors."""
    from ml4ca_b200 import tf_checkpoint
    base = os.path.join(ref_loader.REFERENCE_ROOT, 'src', 'rl')
    models = {
        'final_80x3': os.path.join(base, 'windows_workspace', 'data', 'finalmodel', 'finconttothighbowder_s0', 'tf1_save'),
        'limited_64x3': os.path.join(base, 'ROS', 'rl_allocator', 'src', 'models', 'limited', 'tf1_save'),
    }
    for tag, path in models.items():
        flat, dims = tf_checkpoint.load_actor_critic(path)
        np.savez_compressed(os.path.join(HERE, 'policy_%s.npz' % tag), params=flat,
                            **{k: np.array(v) for k, v in dims.items()}, activation=np.array('leaky_relu'),
                            source=np.array(os.path.relpath(path, ref_loader.REFERENCE_ROOT)))
        print('wrote policy_%s.npz' % tag, dims, flat.shape)


def gen_ros_adapter(K=96, seed=21):
    """The deployment node itself (src/rl/ROS/rl_allocator/src/rl_allocator.py, class RLTA) driven message by message:
    eta / nu / desired-state callbacks with a stub actor returning a prescribed network action.  Records the state the
    actor saw, the ROS-order action u, the published message fields and the state tail after the step."""
    import sys as _sys
    mod = ref_loader.load_rl_allocator_module()
    mod.load_policy = lambda fpath, num_hidden_layers=None: (lambda s: np.zeros(7))
    Twist = _sys.modules["geometry_msgs.msg"].Twist
    NEH = _sys.modules["custom_msgs.msg"].NorthEastHeading
    rng = np.random.default_rng(seed)
    out = {}
    for tag, env, cont, act_dim in (("final_cont", "final", True, 7), ("final_wrap", "final", False, 5),
                                    ("limited", "limited", False, 5), ("full", "full", False, 6)):
        node = mod.RLTA()
        node.env, node.cont_ang = env, cont
        rec = {k: [] for k in ("eta_deg", "nu", "ref_deg", "action", "state_seen", "u", "msg", "state_after", "prev_u")}
        for k in range(K):
            eta = rng.uniform(-1, 1, 3) * np.array([8.0, 8.0, 400.0])          # heading in degrees, beyond +-180 too
            nu = rng.uniform(-1, 1, 3) * np.array([1.4, 0.3, 0.52])
            ref = rng.uniform(-1, 1, 3) * np.array([8.0, 8.0, 200.0])
            a = rng.uniform(-1.4, 1.4, act_dim)
            seen = {}

            def actor(state, a=a, seen=seen):
                seen["s"] = np.array(state, dtype=np.float64)
                return a.copy()
            node.actor = actor
            rec["prev_u"].append(np.array(node.prev_thrust_state, dtype=np.float64))
            node.eta_obs_callback(Twist(eta[0], eta[1], eta[2]))
            node.nu_obs_callback(Twist(nu[0], nu[1], nu[2]))
            d = NEH()
            d.pos_north, d.pos_east, d.pos_heading = ref
            node.state_desired_callback(d)
            pa, st, bc = node.pub_stern_angles.last, node.pub_stern_thruster_setpoints.last, node.pub_bow_control.last
            rec["eta_deg"].append(eta); rec["nu"].append(nu); rec["ref_deg"].append(ref); rec["action"].append(a)
            rec["state_seen"].append(seen["s"]); rec["u"].append(np.array(node.prev_thrust_state, dtype=np.float64))
            rec["msg"].append([pa.port, pa.star, st.port_effort, st.star_effort, bc.throttle_bow, bc.position_bow, bc.lin_act_bow])
            rec["state_after"].append(np.array(node.state, dtype=np.float64))
        for k, v in rec.items():
            out["%s__%s" % (tag, k)] = np.array(v, dtype=np.float64)
    return out


def gen_reset_acts(B=64, T=3, seed=31):
    """RevoltFinal(reset_acts=True) itself (customEnv.py:179-188): the np.random.normal draws of each reset are
    reproduced beforehand from the same global seed (with explicit **init the reset draws nothing else), then the
    reference is reset and stepped.  Records the standard normals, the reset observation and T steps."""
    mod = ref_loader.load_env_module()
    rng = np.random.default_rng(seed)
    rec = {k: [] for k in ('eta0', 'nu0', 'z', 'obs0', 'actions', 'obs', 'rew')}
    for b in range(B):
        twin = vessel.VesselTwin(frozen=False)
        env = mod.RevoltFinal(twin, extended_state=True, cont_ang=True, reset_acts=True)
        eta0 = rng.uniform(-1, 1, 3) * np.array([6.0, 6.0, 0.6])
        nu0 = rng.uniform(-1, 1, 3) * np.array([0.3, 0.07, 0.12])
        init = {'Hull.PosNED': [eta0[0], eta0[1]], 'Hull.PosAttitude': [0, 0, eta0[2]],
                'Hull.VelocityNu': [nu0[0], nu0[1], 0, 0, 0, nu0[2]]}
        np.random.seed(1000 + b)
        draws = np.random.normal(loc=0.0, scale=0.1, size=3)
        np.random.seed(1000 + b)
        real_normal = np.random.normal
        if b % 16 == 5:                          # a few draws beyond the clip (|100 * draw| > 100): inflate x120
            np.random.normal = lambda loc=0.0, scale=1.0, size=None: real_normal(loc=loc, scale=scale, size=size) * 120.0
            draws = draws * 120.0
        try:
            o0 = env.reset(**init)
        finally:
            np.random.normal = real_normal
        acts, obs, rew = [], [], []
        for t in range(T):
            a = rng.uniform(-1.1, 1.1, env.num_actions)
            o, r, d, _ = env.step(a)
            acts.append(a); obs.append(np.array(o, dtype=np.float64)); rew.append(float(np.asarray(r).ravel()[0]))
        rec['eta0'].append(eta0); rec['nu0'].append(nu0); rec['z'].append(draws / 0.1)
        rec['obs0'].append(np.array(o0, dtype=np.float64)); rec['actions'].append(acts); rec['obs'].append(obs)
        rec['rew'].append(rew)
    return {'eta0': np.array(rec['eta0']).T, 'nu0': np.array(rec['nu0']).T, 'z': np.array(rec['z']).T,
            'obs0': np.array(rec['obs0']).T, 'actions': np.transpose(np.array(rec['actions']), (1, 2, 0)),
            'obs': np.transpose(np.array(rec['obs']), (1, 2, 0)), 'rew': np.array(rec['rew']).T}


def gen_hull_replay(window_s=10.0, pre_s=5.0, every=2):
    """Recorded Cybersea box tests of the reference (results/all_plots/box_test/bagfile__*.csv: observer pose at 20 Hz and
    the thruster commands of four controllers), cut into open-loop replay windows: the recorded pose of the window, the
    body velocity at its start (Savitzky-Golay rates of the pose, tools/sysid_hull.py) and the COMMANDED thruster wrench
    (reference thruster model on the recorded commands) from pre_s before the window, so that an actuator lag can settle."""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    argv, sys.argv = sys.argv, [sys.argv[0]]
    import sysid_hull as S
    sys.argv = argv
    eta_w, nu0, tau_w, src = [], [], [], []
    for k, m in enumerate(('RL', 'QP', 'pseudo', 'RLintegral')):
        t, eta, n, a = S.load(m)
        h = t[1] - t[0]
        W, P = int(round(window_s / h)), int(round(pre_s / h))
        nu, _ = S.body_rates(t, eta)
        tau = S.wrench(n, a)
        for s in list(range(100, len(t) - W - 100, W))[::every]:
            eta_w.append(eta[:, s:s + W]); nu0.append(nu[:, s]); tau_w.append(tau[:, s - P:s + W]); src.append(k)
    return {'eta': np.array(eta_w, dtype=np.float32), 'nu0': np.array(nu0, dtype=np.float32),
            'tau_cmd': np.array(tau_w, dtype=np.float32), 'run': np.array(src, dtype=np.int8),
            'h': np.float64(0.05), 'pre': np.int32(int(round(pre_s / 0.05)))}


def main():
    assert ref_loader.available(), "reference checkout not found"
    if "--only-hull-replay" in sys.argv:
        out = gen_hull_replay()
        np.savez_compressed(os.path.join(HERE, 'hull_replay.npz'), **out)
        print('wrote hull_replay.npz  windows %d' % out['eta'].shape[0])
        return
    if "--only-reset-acts" in sys.argv:
        np.savez_compressed(os.path.join(HERE, 'resetacts_final.npz'), **gen_reset_acts())
        print('wrote resetacts_final.npz')
        return
    if "--only-qp" in sys.argv:
        out = gen_qp(4096, seed=0)
        np.savez_compressed(os.path.join(HERE, 'qp_config1.npz'), **out)
        print('wrote qp_config1.npz  success rate %.4f' % out['success'].mean())
        return
    if "--only-error-frame" in sys.argv:
        np.savez_compressed(os.path.join(HERE, 'error_frame.npz'), **gen_error_frame())
        print('wrote error_frame.npz')
        return
    if "--only-gae" in sys.argv:
        np.savez_compressed(os.path.join(HERE, 'gae.npz'), **gen_gae(7))
        print('wrote gae.npz')
        return
    if "--only-qp-switches" in sys.argv:
        np.savez_compressed(os.path.join(HERE, 'qp_switches.npz'), **gen_qp_switches())
        print('wrote qp_switches.npz')
        return
    if "--only-ros" in sys.argv:
        np.savez_compressed(os.path.join(HERE, 'ros_adapter.npz'), **gen_ros_adapter())
        print('wrote ros_adapter.npz')
        return
    cases = [
        ('final_cont_ext', 'final', 'RevoltFinal', dict(cont_ang=True, extended_state=True), 32, 40),
        ('final_wrap_ext', 'final', 'RevoltFinal', dict(cont_ang=False, extended_state=True), 8, 20),
        ('final_cont_std', 'final', 'RevoltFinal', dict(cont_ang=True, extended_state=False), 8, 20),
        ('limited_ext', 'limited', 'RevoltLimited', dict(extended_state=True), 8, 20),
        ('simple_std', 'simple', 'RevoltSimple', dict(extended_state=False), 8, 20),
        ('full_ext', 'full', 'Revolt', dict(extended_state=True), 8, 20),
    ]
    for i, (tag, kind, cls, kw, B, T) in enumerate(cases):
        for mode, frozen in (('null', True), ('hull', False)):
            out = gen_env(kind, cls, kw, frozen, B, T, seed=100 + i)
            out['kind'] = np.array(kind)
            out['cont_ang'] = np.array(bool(kw.get('cont_ang', False)))
            out['extended_state'] = np.array(bool(kw['extended_state']))
            path = os.path.join(HERE, 'env_%s_%s.npz' % (tag, mode))
            np.savez_compressed(path, **out)
            print('wrote', path)
    out = gen_qp(4096, seed=0)    # BASELINE configs[0] / SURVEY 8(d): the full 4096-demand batch
    np.savez_compressed(os.path.join(HERE, 'qp_config1.npz'), **out)
    print('wrote qp_config1.npz  success rate %.3f' % out['success'].mean())
    np.savez_compressed(os.path.join(HERE, 'qp_switches.npz'), **gen_qp_switches())
    print('wrote qp_switches.npz')
    np.savez_compressed(os.path.join(HERE, 'resetacts_final.npz'), **gen_reset_acts())
    print('wrote resetacts_final.npz')
    np.savez_compressed(os.path.join(HERE, 'gae.npz'), **gen_gae(7))
    print('wrote gae.npz')
    np.savez_compressed(os.path.join(HERE, 'error_frame.npz'), **gen_error_frame())
    print('wrote error_frame.npz')
    gen_policy()
    np.savez_compressed(os.path.join(HERE, 'ros_adapter.npz'), **gen_ros_adapter())
    print('wrote ros_adapter.npz')


if __name__ == '__main__':
    main()
