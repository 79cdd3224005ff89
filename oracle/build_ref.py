"""Recipe for oracle/_ref/: the reference's OWN implementation of the path, compiled where its sources lie.

TEST INFRASTRUCTURE (see oracle/__init__.py).  The reference is Python, so "compiling" means byte-compiling: the few
modules on the hot path are compiled by ``py_compile`` straight from /root/reference into sourceless byte-code files under
oracle/_ref/ (git-ignored, but it travels to the GPU box with the repository snapshot like the built .so; the files carry
the extension ``.refbin`` because the snapshot tool drops ``*.pyc``, and ``oracle.ref_loader`` imports them through
``importlib.machinery.SourcelessFileLoader``).  No reference
source text is copied into the repository.  With oracle/_ref/ present, ``oracle.ref_loader`` can import the reference's
customEnv (Revolt*, ErrorFrame, mathematics, simtools) and qp_allocator (QPTA) on a box that has no /root/reference, which
is what lets ``bench.py --impl reference`` and the ``cpu_baseline`` leg time the REFERENCE CODE (kind "reference") rather
than the NumPy port.  Same interpreter on both sides (the image is identical), so the .pyc files load.

    python -m oracle.build_ref            # or __graft_entry__.build(), which calls build() when /root/reference exists
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("ML4CA_REFERENCE_ROOT", "/root/reference")

# (source relative to the reference root, module path under oracle/_ref/)
MODULES = [
    ("src/rl/windows_workspace/specific/customEnv.py", "specific.customEnv.refbin"),
    ("src/rl/windows_workspace/specific/errorFrame.py", "specific.errorFrame.refbin"),
    ("src/rl/windows_workspace/specific/misc/mathematics.py", "specific.misc.mathematics.refbin"),
    ("src/rl/windows_workspace/specific/misc/simtools.py", "specific.misc.simtools.refbin"),
    ("src/qp/ROS/qp_allocator/src/qp_allocator.py", "qp_allocator.refbin"),
]


def build(verbose=False):
    """Returns the list of compiled files (empty when the reference checkout is absent: nothing to do)."""
    if not os.path.isdir(REFERENCE_ROOT):
        return []
    done = []
    for src, dst in MODULES:
        s, d = os.path.join(REFERENCE_ROOT, src), os.path.join(OUT, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        py_compile.compile(s, cfile=d, doraise=True, optimize=0)
        done.append(d)
        if verbose:
            print("compiled %s -> %s" % (src, os.path.relpath(d, HERE)))
    with open(os.path.join(OUT, "README"), "w") as fh:
        fh.write("byte-compiled from %s by oracle/build_ref.py with %s; test infrastructure, not product\n"
                 % (REFERENCE_ROOT, sys.version.split()[0]))
    return done


if __name__ == "__main__":
    print("\n".join(build(verbose=True)) or "reference checkout not present: nothing built")
