"""CPU oracle for the ReVolt dynamic-positioning hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ml4ca_b200/`` may import this package; the only
legitimate importers are ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs.  It is the checker, never the thing shipped or measured as the product.

Contents
--------
constants.py   numbers mirrored from ``ml4ca_b200/csrc/ml4ca_constants.h`` (kept equal by a test)
philox.py      Philox4x32-10 counter RNG in NumPy (bit-identical to csrc/philox.cuh)
vessel.py      float64 NumPy integration of the DECLARED stand-in 3-DOF hull equations
env_oracle.py  vectorised NumPy restatement of specific/customEnv.py (RevoltFinal & friends)
qp_oracle.py   the reference NLP of qp_allocator.py::solve_QP restated on SciPy SLSQP + a tight solve
pinv_oracle.py float64 restatement of this build's stated pseudoinverse + PID equations
mlp_oracle.py  NumPy restatement of spinup/algos/tf1/ppo/core.py (MLP, Gaussian policy, logp)
ppo_oracle.py  GAE / discounted cumsum / PPO losses restated (ppo.py:65-105,234-249)
ros_oracle.py  state assembly / action post-processing of the deployment node (rl_allocator.py, utils.py)
eval_oracle.py IAE / W* / IADC of results/all_plots (common.py, box_test/plot_act.py)
ref_loader.py  imports the UNMODIFIED reference modules from /root/reference behind stub modules
               (only usable in the build container; used to pin the restatements and to
               generate tests/golden/*.npz)

Parity pinning status (also stated in DESIGN.md):
  env / obs / reward / termination .... pinned against the reference code itself (golden vectors)
  QP allocator ........................ pinned against the reference code + container SciPy 1.18.1
                                        (reference pinned scipy==1.2.0: that solver build is absent)
  GAE / discount_cumsum ............... pinned against the reference's scipy.signal formulation
  MLP forward ......................... restatement only (TensorFlow 1 not installable) - weights from
                                        the shipped checkpoints are used as fixtures
  pseudoinverse + PID, hull dynamics .. PARITY UNPINNED: absent from the reference, equations declared
"""
