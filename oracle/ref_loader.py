"""Import the UNMODIFIED reference modules from /root/reference behind stub modules.

TEST INFRASTRUCTURE (see oracle/__init__.py).  In the build container the modules come from the
read-only checkout; on the GPU box, where /root/reference does not exist, the env wrapper and the QP
allocator come from oracle/_ref/ (byte-compiled from the checkout by oracle/build_ref.py, travels with the
repository snapshot), everything else from the committed golden vectors under tests/golden/.

Stubbed third-party modules (absent here, and irrelevant to the arithmetic on the hot path):
  gym / gym.spaces   -> ``Env`` base class and a ``Box`` record   (customEnv.py:1-2,62-64)
  keras / keras.backend -> dead import in specific/misc/mathematics.py:2
  rospy, custom_msgs.msg, geometry_msgs.msg -> ROS I/O of qp_allocator.py:17-19,40-87
``rospy.get_time()`` is pinned to 0.0, which disables the wall-clock keyed retry loop of
qp_allocator.py:209 and makes ``solve_QP`` deterministic.
"""
import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ML4CA_REFERENCE_ROOT", "/root/reference")
_RL_ROOT = os.path.join(REFERENCE_ROOT, "src", "rl", "windows_workspace")
_QP_ROOT = os.path.join(REFERENCE_ROOT, "src", "qp", "ROS", "qp_allocator", "src")
# oracle/_ref/: the same modules byte-compiled from the checkout by oracle/build_ref.py (travels to the GPU box)
_REF_COMPILED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def have_checkout():
    return os.path.isfile(os.path.join(_RL_ROOT, "specific", "customEnv.py"))


_COMPILED_MODULES = ("specific.customEnv", "specific.errorFrame", "specific.misc.mathematics", "specific.misc.simtools",
                     "qp_allocator")


def have_compiled():
    return all(os.path.isfile(os.path.join(_REF_COMPILED, m + ".refbin")) for m in _COMPILED_MODULES)


class _CompiledFinder(object):
    """Meta-path finder for oracle/_ref/<dotted name>.refbin (sourceless byte code of the reference modules; `specific` and
    `specific.misc` are empty namespace packages, as in the reference)."""

    @staticmethod
    def find_spec(name, path=None, target=None):
        import importlib.machinery as mach
        if name in ("specific", "specific.misc"):
            spec = mach.ModuleSpec(name, None, is_package=True)
            spec.submodule_search_locations = []
            return spec
        if name in _COMPILED_MODULES:
            f = os.path.join(_REF_COMPILED, name + ".refbin")
            return importlib.util.spec_from_file_location(name, f, loader=mach.SourcelessFileLoader(name, f))
        return None


def available():
    """The reference's env wrapper and QP allocator can be imported (from the checkout, else from oracle/_ref)."""
    return have_checkout() or have_compiled()


def source():
    return "checkout" if have_checkout() else ("oracle/_ref" if have_compiled() else None)


def _prepare_import():
    """Make `specific.*` and `qp_allocator` importable from the checkout, else from oracle/_ref."""
    if have_checkout():
        for root in (_RL_ROOT, _QP_ROOT):
            if root not in sys.path:
                sys.path.insert(0, root)
    elif not any(isinstance(f, type) and f is _CompiledFinder for f in sys.meta_path):
        sys.meta_path.insert(0, _CompiledFinder)


def _install_stubs():
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env(object):
            pass

        class Box(object):
            def __init__(self, low=None, high=None, dtype=None, shape=None):
                self.low, self.high, self.dtype = low, high, dtype
                self.shape = tuple(low.shape) if shape is None else tuple(shape)

        spaces = types.ModuleType("gym.spaces")
        spaces.Box = Box
        spaces.Discrete = type("Discrete", (), {})
        gym.Env = Env
        gym.spaces = spaces
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    if "keras" not in sys.modules:
        keras = types.ModuleType("keras")
        backend = types.ModuleType("keras.backend")
        keras.backend = backend
        sys.modules["keras"] = keras
        sys.modules["keras.backend"] = backend
    if "rospy" not in sys.modules:
        rospy = types.ModuleType("rospy")
        rospy.init_node = lambda *a, **k: None
        rospy.Rate = lambda hz: types.SimpleNamespace(sleep=lambda: None)
        rospy.get_time = lambda: 0.0
        rospy.loginfo = lambda *a, **k: None
        rospy.logwarn = lambda *a, **k: None
        rospy.spin = lambda: None
        rospy.ROSInterruptException = type("ROSInterruptException", (Exception,), {})

        class _Pub(object):
            def __init__(self, *a, **k):
                self.last = None

            def publish(self, msg):
                self.last = msg

        rospy.Publisher = _Pub
        rospy.Subscriber = lambda *a, **k: None
        sys.modules["rospy"] = rospy

        def _msg(name, fields):
            def __init__(self):
                for f in fields:
                    setattr(self, f, 0.0)
            return type(name, (), {"__init__": __init__})

        custom = types.ModuleType("custom_msgs")
        cmsg = types.ModuleType("custom_msgs.msg")
        cmsg.podAngle = _msg("podAngle", ["port", "star"])
        cmsg.SternThrusterSetpoints = _msg("SternThrusterSetpoints", ["port_effort", "star_effort"])
        cmsg.bowControl = _msg("bowControl", ["throttle_bow", "position_bow", "lin_act_bow"])
        cmsg.diffThrottleStern = _msg("diffThrottleStern", ["throttle", "rudder"])
        custom.msg = cmsg
        sys.modules["custom_msgs"] = custom
        sys.modules["custom_msgs.msg"] = cmsg

        geo = types.ModuleType("geometry_msgs")
        gmsg = types.ModuleType("geometry_msgs.msg")

        class Wrench(object):
            def __init__(self, fx=0.0, fy=0.0, tz=0.0):
                self.force = types.SimpleNamespace(x=fx, y=fy, z=0.0)
                self.torque = types.SimpleNamespace(x=0.0, y=0.0, z=tz)

        gmsg.Wrench = Wrench
        geo.msg = gmsg
        sys.modules["geometry_msgs"] = geo
        sys.modules["geometry_msgs.msg"] = gmsg


def load_env_module():
    """-> the reference ``specific.customEnv`` module (Revolt, RevoltFinal, ...)."""
    if not available():
        raise RuntimeError("reference not present (neither %s nor oracle/_ref)" % REFERENCE_ROOT)
    _install_stubs()
    _prepare_import()
    return importlib.import_module("specific.customEnv")


def load_error_frame_module():
    load_env_module()
    return importlib.import_module("specific.errorFrame")


def load_qp_module():
    """-> the reference ROS ``qp_allocator`` module (class QPTA)."""
    if not available():
        raise RuntimeError("reference not present (neither %s nor oracle/_ref)" % REFERENCE_ROOT)
    _install_stubs()
    _prepare_import()
    return importlib.import_module("qp_allocator")


def load_rl_allocator_module():
    """-> the reference deployment node module ``rl_allocator`` (class RLTA) with its siblings ``errorFrame`` and
    ``utils`` (src/rl/ROS/rl_allocator/src).  TensorFlow, ROS messages and the policy loader are stubbed; construct
    ``RLTA()`` after setting ``module.load_policy`` to a function returning the actor stub."""
    if not have_checkout():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    _install_stubs()
    cmsg = sys.modules["custom_msgs.msg"]
    if not hasattr(cmsg, "NorthEastHeading"):
        fields = ["pos_north", "pos_east", "pos_heading", "vel_north", "vel_east", "vel_heading"]

        def _init(self):
            for f in fields:
                setattr(self, f, 0.0)
        cmsg.NorthEastHeading = type("NorthEastHeading", (), {"__init__": _init})
    gmsg = sys.modules["geometry_msgs.msg"]
    if not hasattr(gmsg, "Twist"):
        class Twist(object):
            def __init__(self, x=0.0, y=0.0, az=0.0):
                self.linear = types.SimpleNamespace(x=x, y=y, z=0.0)
                self.angular = types.SimpleNamespace(x=0.0, y=0.0, z=az)
        gmsg.Twist = Twist
        gmsg.Pose2D = type("Pose2D", (), {})
    if "std_msgs" not in sys.modules:
        std = types.ModuleType("std_msgs")
        smsg = types.ModuleType("std_msgs.msg")
        smsg.Float64 = type("Float64", (), {})
        std.msg = smsg
        sys.modules["std_msgs"] = std
        sys.modules["std_msgs.msg"] = smsg
    if "tensorflow" not in sys.modules:
        sys.modules["tensorflow"] = types.ModuleType("tensorflow")     # utils.py:7, only used by the (stubbed) loader
    rospy = sys.modules["rospy"]
    if not hasattr(rospy, "logerr"):
        rospy.logerr = lambda *a, **k: None
        rospy.on_shutdown = lambda *a, **k: None
    root = os.path.join(REFERENCE_ROOT, "src", "rl", "ROS", "rl_allocator", "src")
    saved = {k: sys.modules.get(k) for k in ("errorFrame", "utils")}
    try:
        for name in ("errorFrame", "utils", "rl_allocator"):
            spec = importlib.util.spec_from_file_location(name, os.path.join(root, name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
        return sys.modules["rl_allocator"]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def load_ppo_module():
    """-> the reference ``spinup.algos.tf1.ppo.ppo`` module (TrajectoryBuffer, and through it core.discount_cumsum and
    mpi_tools.mpi_statistics_scalar), imported UNMODIFIED behind stubs for tensorflow (only touched inside functions that
    build graphs, and for default arguments), gym and mpi4py (one process: Allreduce copies).  Checkout only."""
    if not have_checkout():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    _install_stubs()

    class _Anything(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            child = _Anything(self.__name__ + "." + name)
            setattr(self, name, child)
            return child

        def __call__(self, *a, **k):
            return self

    if "tensorflow" not in sys.modules or not isinstance(sys.modules["tensorflow"], _Anything):
        tf = _Anything("tensorflow")
        tf.train.AdamOptimizer = type("AdamOptimizer", (object,), {})    # base class of mpi_tf.MpiAdamOptimizer (:29)
        sys.modules["tensorflow"] = tf
    gym_spaces = sys.modules["gym.spaces"]
    if "mpi4py" not in sys.modules:
        import numpy as np
        mpi4py = types.ModuleType("mpi4py")
        MPI = types.ModuleType("mpi4py.MPI")

        class _Comm(object):
            def Get_rank(self):
                return 0

            def Get_size(self):
                return 1

            def Allreduce(self, sendbuf, recvbuf, op=None):
                np.copyto(recvbuf, sendbuf)

            def Bcast(self, x, root=0):
                pass

        MPI.COMM_WORLD = _Comm()
        MPI.SUM, MPI.MIN, MPI.MAX = "sum", "min", "max"
        mpi4py.MPI = MPI
        sys.modules["mpi4py"] = mpi4py
        sys.modules["mpi4py.MPI"] = MPI
    assert hasattr(gym_spaces, "Box") and hasattr(gym_spaces, "Discrete")
    if _RL_ROOT not in sys.path:
        sys.path.insert(0, _RL_ROOT)
    return importlib.import_module("spinup.algos.tf1.ppo.ppo")


def wrench(fx, fy, tz):
    _install_stubs()
    return sys.modules["geometry_msgs.msg"].Wrench(fx, fy, tz)
