"""the non-Gaussian targets of the BASELINE configurations on the device against the outputs of
the reference's own classes (tests/golden/reference_targets.npz, tests/golden/make_reference_targets.py) and against the
oracle restatement.  Tolerances: log-densities <= 1e-5, gradients <= 1e-4 (relative to the largest magnitude)."""
import os

import numpy as np
import pytest
import torch

import oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_targets.npz")


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


@pytest.mark.parametrize("name", ["stm20", "stm_hard64"])
def test_student_t_mixture_matches_reference_class(name):
    """StudentTMixture_LNPDF.log_density (student_t_mixture.py:34-68) + the tape gradient of sample_selector.py:73-77."""
    from gmmvi_b200.experiments.target_distributions.student_t_mixture import StudentTMixture_LNPDF
    g = np.load(GOLD)
    ch = g[name + "_chols"]
    K = ch.shape[0]
    tgt = StudentTMixture_LNPDF(np.ones(K) / K, g[name + "_means"], ch @ ch.transpose(0, 2, 1))
    X = torch.as_tensor(g[name + "_X"], dtype=torch.float32).cuda()
    lp = tgt.log_density(X)
    lp2, grad = tgt.log_density_and_grad(X)
    assert torch.equal(lp, lp2)
    e_v, e_g = rel_err(lp.cpu().numpy(), g[name + "_lnpdf"]), rel_err(grad.cpu().numpy(), g[name + "_grad"])
    # element-wise on the log-density: every sample, not only the largest magnitude
    e_el = float(np.max(np.abs(lp.cpu().numpy() - g[name + "_lnpdf"]) / np.maximum(np.abs(g[name + "_lnpdf"]), 1.0)))
    print(f"{name}: lnpdf {e_v:.2e} (element-wise {e_el:.2e}) grad {e_g:.2e}")
    assert e_v < 1e-5 and e_el < 1e-5 and e_g < 1e-4


@pytest.mark.parametrize("name,links,goals", [("planar10_4", 10, 4), ("planar10_1", 10, 1), ("planar3_4", 3, 4)])
def test_planar_robot_matches_reference_class(name, links, goals):
    """PlanarRobot.log_density / forward_kinematics (planar_robot.py:49-66) + the tape gradient."""
    from gmmvi_b200.experiments.target_distributions.planar_robot import PlanarRobot
    g = np.load(GOLD)
    tgt = PlanarRobot(links, goals)
    X = torch.as_tensor(g[name + "_X"], dtype=torch.float32).cuda()
    lp, grad = tgt.log_density_and_grad(X)
    assert torch.equal(tgt.log_density(X), lp)
    e_v, e_g = rel_err(lp.cpu().numpy(), g[name + "_lnpdf"]), rel_err(grad.cpu().numpy(), g[name + "_grad"])
    e_fk = rel_err(tgt.forward_kinematics(X).cpu().numpy(), g[name + "_fk"])
    print(f"{name}: lnpdf {e_v:.2e} grad {e_g:.2e} fk {e_fk:.2e}")
    assert e_v < 1e-5 and e_g < 1e-4 and e_fk < 1e-5


def test_planar_robot_near_the_goal_matches_oracle():
    """Samples whose end effector is within a few likelihood standard deviations of a goal (where the posterior mass is):
    the difference pos - goal cancels, so the check is against the fp64 oracle with the fp32 noise floor of the reference's
    own formula beside it."""
    from gmmvi_b200.experiments.target_distributions.planar_robot import PlanarRobot
    rng = np.random.default_rng(5)
    base = np.zeros(10)
    base[0] = 0.0
    # straight arm of length 10 reaches x = 10; fold it to reach (7, 0): theta = (a, -2a, 2a, -2a, ...) zig-zag
    a = np.arccos(0.7)
    zig = np.array([a] + [(-2 * a) if i % 2 else (2 * a) for i in range(1, 10)])
    X = (zig[None, :] + 1e-3 * rng.standard_normal((512, 10))).astype(np.float32)
    ref_v, ref_g = O.planar_robot_target(10, 4)(X.astype(np.float64))
    f32_v, f32_g = O.planar_robot_target(10, 4, dt=np.float32)(X)
    lp, grad = PlanarRobot(10, 4).log_density_and_grad(torch.as_tensor(X).cuda())
    e_v, e_g = rel_err(lp.cpu().numpy(), ref_v), rel_err(grad.cpu().numpy(), ref_g)
    n_v, n_g = rel_err(f32_v, ref_v), rel_err(f32_g, ref_g)
    print(f"near goal: lnpdf {e_v:.2e} (fp32 formula {n_v:.2e}) grad {e_g:.2e} (fp32 formula {n_g:.2e})")
    assert e_v < max(1e-5, 3 * n_v) and e_g < max(1e-4, 3 * n_g)


def test_make_target_runs_through_the_selector():
    """The targets plug into SampleSelector.get_target_grads (sample_selector.py:69-78)."""
    from gmmvi_b200.experiments.target_distributions.planar_robot import make_four_goal
    from gmmvi_b200.experiments.target_distributions.student_t_mixture import make_target
    from gmmvi_b200.optimization.gmmvi_modules.sample_selector import SampleSelector
    for tgt in (make_four_goal(), make_target(20, False)):
        sel = SampleSelector(tgt, None, None)
        X = torch.randn(300, tgt.get_num_dimensions(), device="cuda")
        grad, val = sel.get_target_grads(X)
        assert grad.shape == X.shape and val.shape == (300,)
        assert torch.isfinite(grad).all() and torch.isfinite(val).all()
