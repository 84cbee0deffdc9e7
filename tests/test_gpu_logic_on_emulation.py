"""The bodies of the `-m gpu` tests that were written after the round's last GPU run, executed on the CPU with the kernel emulation in
place of the GPU stepper (tests/hostemu/adapter.py): their Python logic and their thresholds are checked here; on the B200 the same
functions run against libdojo_b200.so.

Emulating 64 environments x 40 steps on CPU fibers takes minutes, so only the shortest case runs by default; the others run with
DOJO_EMULATE_GPU_TESTS=1 (they passed at the end of round 1: the seven short cases in 11 min, the quadruped and ant parity cases in 11 and 9 min more)."""
import os

import pytest

slow = pytest.mark.skipif(not os.environ.get("DOJO_EMULATE_GPU_TESTS"), reason="minutes of CPU time: set DOJO_EMULATE_GPU_TESTS=1")


@pytest.fixture()
def emulated_stepper(monkeypatch):
    import dojo_jl_b200.solver as solver
    import dojo_jl_b200.environments as environments
    from hostemu.adapter import EmuStepper
    monkeypatch.setattr(solver, "BatchedStepper", EmuStepper)
    monkeypatch.setattr(environments, "BatchedStepper", EmuStepper)
    return EmuStepper


@slow
def test_raiberthopper_parity(emulated_stepper):
    import test_zzzz_gpu_translational as G
    G.test_raiberthopper_parity()


def test_recording_with_translational_impulses(emulated_stepper):
    import test_zzzz_gpu_translational as G
    G.test_recording_with_translational_impulses()


@slow
def test_quadruped_waypoint_environment(emulated_stepper):
    import test_zzzz_gpu_translational as G
    G.test_quadruped_waypoint_environment()


@slow
def test_cartpole_environment_and_minimal_gradients(emulated_stepper):
    import test_zzzz_gpu_translational as G
    G.test_cartpole_environment_and_minimal_gradients()


@slow
@pytest.mark.parametrize("case", ["planar"])
def test_step_rollout_and_gradient_parity(emulated_stepper, case):
    import test_zzzz_gpu_translational as G
    G.test_step_rollout_and_gradient_parity(case)


@slow
@pytest.mark.parametrize("name,ct", [("sphere", "linear"), ("block", "impact")])
def test_contact_models_step_and_gradient_parity(emulated_stepper, name, ct):
    import test_zzzz_gpu_contact_models as G
    G.test_step_and_gradient_parity(name, ct)


@slow
@pytest.mark.parametrize("name,B,T,scale", [("quadruped", 64, 30, 2.0), ("ant", 96, 25, 1.0)])
def test_step_parity_of_the_baseline_models(emulated_stepper, name, B, T, scale):
    """tests/test_gpu_parity.py::test_step_parity (ten minutes and more per case on CPU fibers)"""
    import test_gpu_parity as G
    G._compare_rollout(name, B, T, seed=7, scale=scale)
