"""More of the reference's own behaviour / conservation tests restated on the oracle (and, where cheap, on the device kernels through
the emulation): they have KNOWN ANSWERS (a rest state, an analytic velocity, a conserved quantity), which is what pins an oracle that
cannot be compared with Julia output in this image.

  test/behaviors.jl:21-40   "Box toss"            a thrown block comes to rest: |v| < 1e-8, z = 0.25 +- 1e-3, for four time steps
  test/behaviors.jl:42-55   "Box external force"  set_external_force!: v = F t / m, w = tau t / J
  test/energy.jl:98-130     "Dice"                free rigid body under gravity: mechanical energy constant to 1e-8
"""
import numpy as np
import pytest

import dojo_jl_b200 as dj
from dojo_jl_b200 import capi, quat as Q
from oracle.oracle import Oracle


def _block(timestep=0.01, gravity=-9.81, friction=0.8, contact=True):
    m = dj.get_mechanism("block", timestep=timestep, gravity=gravity)
    if not contact:
        m.contacts = []
    for c in m.contacts:
        c.friction = friction
    return m


def _toss_state(m, h):
    z = m.z0.copy()
    z[0:3] = [0.0, 0.0, 0.5 + 0.25]          # initialize_block!: position + z * edge_length / 2 (block/mechanism.jl:76-95)
    z[3:6] = [1.0, 1.5, 1.0]
    z[10:13] = np.array([5.0, 4.0, 2.0]) * h
    return z


@pytest.mark.parametrize("h", [0.10, 0.05, 0.01, 0.005])
def test_box_toss_comes_to_rest(h):
    m = _block(timestep=h, friction=0.1)
    o = Oracle(m, capi.solver_options(rtol=1e-6, btol=1e-6))
    z = _toss_state(m, h)
    for _ in range(int(round(5.0 / h))):
        z, st, _ = o.step(z, np.zeros(m.nu))
        assert st == 0
    assert np.abs(z[3:6]).max() < 1e-8 and abs(z[2] - 0.25) < 1e-3


def test_box_toss_on_the_device_kernels():
    """the same toss through dojo_step_kernel (emulation), h = 0.05: same rest state, same trajectory as the oracle"""
    from hostemu.harness import HostEmu
    h = 0.05
    m = _block(timestep=h, friction=0.1)
    opts = capi.solver_options(rtol=1e-6, btol=1e-6)
    o, em = Oracle(m, opts), HostEmu(m)
    z = _toss_state(m, h)
    zd = z.copy()[None]
    for k in range(100):
        z, st, it = o.step(z, np.zeros(m.nu))
        zd, sd, itd, _ = em.step(zd, np.zeros((1, m.nu)), opts)
        assert (sd[0], itd[0]) == (st, it) and np.abs(zd[0] - z).max() < 1e-9
    assert np.abs(zd[0, 3:6]).max() < 1e-8 and abs(zd[0, 2] - 0.25) < 1e-3


def test_box_external_force_and_torque():
    m = _block(gravity=0.0, contact=False)
    m.bodies[0].inertia = np.eye(3)
    o = Oracle(m)
    q = Q.rot_z(np.pi / 2)
    for force, torque, check in (([1.0, 0, 0], [0.0, 0, 0], lambda z: abs(z[4] - 0.5)), ([0.0, 0, 0], [1.0, 0, 0], lambda z: abs(z[10] - 0.5))):
        z = m.z0.copy()
        z[0:3] = 0.0
        z[3:6] = 0.0
        z[6:10] = q
        z[10:13] = 0.0
        for k in range(1, 101):
            fext = np.zeros(6)
            if k <= 50:  # set_external_force!(body; force, torque, vertex = [0.5, 0, 0]) (bodies/set.jl:110-115)
                fext[0:3] = Q.qrot(np.array(force), z[6:10])
                fext[3:6] = np.array(torque) + np.cross([0.5, 0.0, 0.0], force)
            z, st, _ = o.step(z, np.zeros(m.nu), fext=fext)
            assert st == 0
        assert check(z) < 1e-3


def test_dice_energy_conservation():
    m = _block(gravity=-10.0, contact=False)
    o = Oracle(m, capi.solver_options(rtol=1e-12, btol=1e-12))
    z = m.z0.copy()
    z[3:6] = [1.0, 2.0, 3.0]
    z[10:13] = [1.0, 1.0, 1.0]
    me = []
    for _ in range(500):
        z, st, _ = o.step(z, np.zeros(m.nu))
        _, diag = o.storage_record()
        me.append(diag[6] + diag[7])
    me = np.array(me[100:])
    assert np.abs((me - me[0]) / me.mean()).max() < 1e-8


def test_quadruped_energy_band_with_springs():
    """test/energy.jl "Quadruped" (:427-463): zero gravity, springs = 1 on every joint (parse_springs = parse_dampers = false: no
    dampers), no limits, no contacts, released from initialize_quadruped! away from the springs' rest pose: the mechanical energy
    (kinetic + spring potential, as recorded by save_to_storage!) stays within 1e-2 over 5 s"""
    m = dj.get_mechanism("quadruped", gravity=0.0, springs=1.0)
    m.contacts = []
    for j in m.joints:
        j.tra.damper = j.rot.damper = 0.0
        j.rot.limit_lo = j.rot.limit_hi = None
    o = Oracle(m, capi.solver_options(rtol=1e-12, btol=1e-12))
    z = m.z0.copy()
    me = []
    for _ in range(500):
        z, st, _ = o.step(z, np.zeros(m.nu))
        assert st == 0
        _, diag = o.storage_record()
        me.append(diag[6] + diag[7])
    me = np.array(me[100:])
    assert me.mean() > 1e-3 and np.abs((me - me[0]) / me.mean()).max() < 1e-2


def test_box_and_pendulum_momentum():
    """test/momentum.jl "Box" (:45-67): a free block with v = [1, 2, 3], w = [10, 10, 10] in zero gravity keeps its linear and angular
    momentum to 1e-8; "Pendulum" (:80-103): the angular momentum of a pendulum in zero gravity (about the fixed joint axis, the
    component the reference checks) is constant to 1e-8"""
    m = _block(gravity=0.0, contact=False)
    o = Oracle(m, capi.solver_options(rtol=1e-12, btol=1e-12))
    z = m.z0.copy()
    z[3:6] = [1.0, 2.0, 3.0]
    z[10:13] = [10.0, 10.0, 10.0]
    P = []
    for _ in range(500):
        z, st, _ = o.step(z, np.zeros(m.nu))
        P.append(o.momentum())
    P = np.array(P[5:])
    assert np.abs(P - P[0]).max() < 1e-8
    p = dj.get_mechanism("pendulum", gravity=0.0)
    for j in p.joints:
        j.rot.damper = j.rot.spring = 0.0
    o = Oracle(p, capi.solver_options(rtol=1e-12, btol=1e-12))
    z = p.forward_kinematics({"joint": [0.7]})
    # initialize!(mech, :pendulum; angle = 0.7, angular_velocity = 5): rotation about the joint axis x through the joint at the top
    w = np.array([5.0, 0.0, 0.0])
    zz = z.reshape(-1, 13)
    xj = np.array([0.0, 0.0, 1.1])  # the joint vertex (pendulum/mechanism.jl: parent_vertex = (L + 0.1) z)
    zz[0, 3:6] = np.cross(Q.qrot(w, zz[0, 6:10]), zz[0, 0:3] - xj)
    zz[0, 10:13] = w
    L = []
    for _ in range(500):
        z, st, _ = o.step(z, np.zeros(p.nu))
        assert st == 0
        L.append(z[10])  # body-frame angular velocity about the joint axis; J_xx w_x + m r^2 w_x is the conserved angular momentum
    L = np.array(L[10:])
    assert np.abs(L - L[0]).max() < 1e-8
