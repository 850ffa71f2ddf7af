"""A menagerie-style Franka Panda description (nested default classes with childclass, class= on assets, meshes
named after their files, STL + OBJ assets under meshdir, visual geoms with contype 0, fullinertia, autolimits, a split
finger tendon driving one general actuator with affine bias, a joint equality, a keyframe with ctrl) must load UNMODIFIED
-- SURVEY 8(f) N4: "lets real menagerie Panda / Shadow Hand XMLs load unmodified instead of primitive stand-ins".  The
file follows the public structure of mujoco_menagerie/franka_emika_panda/panda.xml (no network here: the mesh assets
are small generated bricks, the tree is shortened to four joints); the reference loads such files through
mj_loadXML (mujoco_env.cpp:771-911).  CPU part: compile facts.  GPU part: forward fields and state-injected steps
against the oracle under the file's own implicitfast integrator."""
import struct

import numpy as np
import pytest

PANDA_XML = r'''<mujoco model="panda">
  <compiler angle="radian" meshdir="assets" autolimits="true"/>
  <option integrator="implicitfast"/>
  <default>
    <default class="panda">
      <material specular="0.5" shininess="0.25"/>
      <joint armature="0.1" damping="1" axis="0 0 1" range="-2.8973 2.8973"/>
      <general dyntype="none" biastype="affine" ctrlrange="-2.8973 2.8973" forcerange="-87 87"/>
      <default class="finger">
        <joint axis="0 1 0" type="slide" range="0 0.04"/>
      </default>
      <default class="visual">
        <geom type="mesh" contype="0" conaffinity="0" group="2"/>
      </default>
      <default class="collision">
        <geom type="mesh" group="3"/>
        <default class="fingertip_pad_collision_1">
          <geom type="box" size="0.0085 0.004 0.0085" pos="0 0.0055 0.0445"/>
        </default>
      </default>
    </default>
  </default>
  <asset>
    <material class="panda" name="white" rgba="1 1 1 1"/>
    <material class="panda" name="off_white" rgba="0.901961 0.921569 0.929412 1"/>
    <mesh name="link0_c" file="link0.stl"/>
    <mesh name="link1_c" file="link1.stl"/>
    <mesh name="hand_c" file="hand.stl"/>
    <mesh file="link0_0.obj"/>
    <mesh file="link1_0.obj"/>
    <mesh file="hand_0.obj"/>
    <mesh file="finger_0.obj"/>
  </asset>
  <worldbody>
    <light name="top" pos="0 0 2" mode="trackcom"/>
    <geom name="floor" type="plane" size="0 0 0.05"/>
    <body name="link0" childclass="panda">
      <inertial mass="0.629769" pos="-0.041018 -0.00014 0.049974"
        fullinertia="0.00315 0.00388 0.004285 8.2904e-7 0.00015 8.2299e-6"/>
      <geom mesh="link0_0" material="off_white" class="visual"/>
      <geom mesh="link0_c" class="collision"/>
      <body name="link1" pos="0 0 0.333">
        <inertial mass="4.970684" pos="0.003875 0.002081 -0.04762"
          fullinertia="0.70337 0.70661 0.0091170 -0.00013900 0.0067720 0.019169"/>
        <joint name="joint1"/>
        <geom material="white" mesh="link1_0" class="visual"/>
        <geom mesh="link1_c" class="collision"/>
        <body name="link2" quat="1 -1 0 0">
          <inertial mass="0.646926" pos="-0.003141 -0.02872 0.003495"
            fullinertia="0.0079620 2.8110e-2 2.5995e-2 -3.925e-3 1.0254e-2 7.04e-4"/>
          <joint name="joint2" range="-1.7628 1.7628"/>
          <geom type="capsule" size="0.04 0.1" class="collision"/>
          <body name="hand" pos="0 -0.3 0" quat="0.9238795 0 0 -0.3826834">
            <inertial mass="0.73" pos="-0.01 0 0.03" diaginertia="0.001 0.0025 0.0017"/>
            <geom mesh="hand_0" material="off_white" class="visual"/>
            <geom mesh="hand_c" class="collision"/>
            <body name="left_finger" pos="0 0 0.0584">
              <inertial mass="0.015" pos="0 0 0" diaginertia="2.375e-6 2.375e-6 7.5e-7"/>
              <joint name="finger_joint1" class="finger"/>
              <geom mesh="finger_0" material="off_white" class="visual"/>
              <geom class="fingertip_pad_collision_1"/>
            </body>
            <body name="right_finger" pos="0 0 0.0584" quat="0 0 0 1">
              <inertial mass="0.015" pos="0 0 0" diaginertia="2.375e-6 2.375e-6 7.5e-7"/>
              <joint name="finger_joint2" class="finger"/>
              <geom mesh="finger_0" material="off_white" class="visual"/>
              <geom class="fingertip_pad_collision_1"/>
            </body>
          </body>
        </body>
      </body>
    </body>
  </worldbody>
  <tendon>
    <fixed name="split">
      <joint joint="finger_joint1" coef="0.5"/>
      <joint joint="finger_joint2" coef="0.5"/>
    </fixed>
  </tendon>
  <equality>
    <joint joint1="finger_joint1" joint2="finger_joint2" solimp="0.95 0.99 0.001" solref="0.005 1"/>
  </equality>
  <actuator>
    <general class="panda" name="actuator1" joint="joint1" gainprm="4500" biasprm="0 -4500 -450"/>
    <general class="panda" name="actuator2" joint="joint2" gainprm="4500" biasprm="0 -4500 -450" ctrlrange="-1.7628 1.7628"/>
    <general class="panda" name="actuator8" tendon="split" forcerange="-100 100" ctrlrange="0 255"
      gainprm="0.01568627451 0 0" biasprm="0 -100 -10"/>
  </actuator>
  <keyframe>
    <key name="home" qpos="0 0 0.04 0.04" ctrl="0 0 255"/>
  </keyframe>
</mujoco>
'''


def _write_assets(d):
    def box(h):
        c = [(sx * h[0], sy * h[1], sz * h[2]) for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
        faces = [(0, 1, 3), (0, 3, 2), (4, 6, 7), (4, 7, 5), (0, 4, 5), (0, 5, 1), (2, 3, 7), (2, 7, 6), (0, 2, 6), (0, 6, 4),
                 (1, 5, 7), (1, 7, 3)]
        return c, faces

    (d / "assets").mkdir(exist_ok=True)
    for n, h in [("link0", (0.08, 0.08, 0.05)), ("link1", (0.05, 0.05, 0.1)), ("hand", (0.03, 0.09, 0.03)),
                 ("finger", (0.01, 0.01, 0.025))]:
        c, faces = box(h)
        with open(d / "assets" / f"{n}.stl", "wb") as f:
            f.write(b"\0" * 80)
            f.write(struct.pack("<I", len(faces)))
            for fa in faces:
                f.write(struct.pack("<3f", 0, 0, 0))
                for i in fa:
                    f.write(struct.pack("<3f", *c[i]))
                f.write(b"\0\0")
        (d / "assets" / f"{n}_0.obj").write_text("# v\n" + "".join(f"v {v[0]} {v[1]} {v[2]}\n" for v in c) +
                                                "".join(f"f {a + 1} {b + 1} {e + 1}\n" for a, b, e in faces))


@pytest.fixture()
def panda(capi, tmp_path):
    _write_assets(tmp_path)
    p = tmp_path / "panda.xml"
    p.write_text(PANDA_XML)
    return capi.Model.from_xml_file(str(p))


def test_menagerie_style_panda_compiles_verbatim(panda):
    m = panda
    assert (m.nq, m.nv, m.nu, m.nbody, m.ngeom) == (4, 4, 3, 7, 12)
    # joint defaults through childclass="panda" and the nested "finger" class; autolimits from the ranges
    np.testing.assert_allclose(m.jnt_range, [[-2.8973, 2.8973], [-1.7628, 1.7628], [0, 0.04], [0, 0.04]])
    assert m.jnt_limited.tolist() == [1, 1, 1, 1]
    assert m.jnt_type.tolist() == [3, 3, 2, 2]
    np.testing.assert_allclose(m.dof_damping, 1.0)
    np.testing.assert_allclose(m.dof_armature, 0.1)
    # <general class="panda">: affine bias, ctrl / force limits from the class unless overridden on the element
    assert m.actuator_biastype.tolist() == [1, 1, 1]
    np.testing.assert_allclose(m.actuator_ctrlrange, [[-2.8973, 2.8973], [-1.7628, 1.7628], [0, 255]])
    np.testing.assert_allclose(m.actuator_forcerange, [[-87, 87], [-87, 87], [-100, 100]])
    assert m.actuator_ctrllimited.tolist() == [1, 1, 1] and m.actuator_forcelimited.tolist() == [1, 1, 1]
    np.testing.assert_allclose(m.actuator_gainprm[:, 0], [4500, 4500, 0.01568627451])
    np.testing.assert_allclose(m.actuator_biasprm[:, :3], [[0, -4500, -450], [0, -4500, -450], [0, -100, -10]])
    assert m.actuator_trntype.tolist() == [0, 0, 3]  # joint, joint, tendon ("split")
    # visual geoms never collide; collision meshes and the pad boxes do
    assert m.geom_contype.tolist() == [1, 0, 1, 0, 1, 1, 0, 1, 0, 1, 0, 1]
    assert m.geom_type.tolist() == [0, 7, 7, 7, 7, 3, 7, 7, 7, 6, 7, 6]
    np.testing.assert_allclose(m.geom_size[9], [0.0085, 0.004, 0.0085])
    # explicit <inertial>: masses as written, fullinertia diagonalised (trace preserved)
    np.testing.assert_allclose(m.body_mass, [0, 0.629769, 4.970684, 0.646926, 0.73, 0.015, 0.015])
    np.testing.assert_allclose(m.body_inertia[1].sum(), 0.00315 + 0.00388 + 0.004285, rtol=1e-12)
    np.testing.assert_allclose(m.body_inertia[4], [0.001, 0.0025, 0.0017])
    # joint equality with its own solref / solimp (three solimp values given: the rest default)
    assert m.eq_type.tolist() == [2]
    np.testing.assert_allclose(m.eq_solref, [[0.005, 1]])
    np.testing.assert_allclose(m.eq_solimp, [[0.95, 0.99, 0.001, 0.5, 2]])
    np.testing.assert_allclose(m.key_qpos.ravel(), [0, 0, 0.04, 0.04])
    np.testing.assert_allclose(m.key_ctrl.ravel(), [0, 0, 255])


@pytest.mark.gpu
def test_menagerie_style_panda_gpu_parity(panda, capi, orc):
    from mujoco_ros_pkgs_b200.batch import BatchSim
    from parity_util import compare_forward_fields, injected_steps, make_oracles, perturbed

    model, nenv = panda, 6
    qpos, qvel = perturbed(model, nenv, seed=5, amp=0.2)
    qpos[:, 2:] = np.clip(qpos[:, 2:], 0.0, 0.04)
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    sim.keep_intermediates(True)
    sim.forward()
    oracles = make_oracles(orc, model, qpos, qvel)
    for o in oracles:
        o.forward()
    compare_forward_fields(capi, model, sim, oracles, skip={"xfrc_applied"}, tag="menagerie-panda")
    sim.keep_intermediates(False)
    worst, _ = injected_steps(model, sim, oracles, 150, np.random.default_rng(3), tag="menagerie-panda")
    assert worst < 1e-5


# ---- a menagerie-style Shadow Hand (public structure of mujoco_menagerie/shadow_hand/right_hand.xml, one finger kept):
# four levels of nested default classes, <position> actuators whose kp / ctrlrange / forcerange come from the joint's class,
# a class on the mesh assets, impratio + elliptic cones, a coupled-joint tendon actuator, a contact exclude, sensors
HAND_XML = r'''<mujoco model="right_shadow_hand">
  <compiler angle="radian" meshdir="assets" autolimits="true"/>
  <option impratio="10" cone="elliptic"/>
  <default>
    <default class="right_hand">
      <mesh scale="1 1 1"/>
      <joint axis="1 0 0" damping="0.05" armature="0.0002" frictionloss="0.01"/>
      <position forcerange="-1 1"/>
      <default class="wrist">
        <joint damping="0.5"/>
        <default class="wrist_y">
          <joint axis="0 1 0" range="-0.523599 0.174533"/>
          <position kp="10" ctrlrange="-0.523599 0.174533" forcerange="-10 10"/>
        </default>
        <default class="wrist_x">
          <joint range="-0.698132 0.488692"/>
          <position kp="8" ctrlrange="-0.698132 0.488692" forcerange="-5 5"/>
        </default>
      </default>
      <default class="knuckle">
        <joint axis="0 -1 0" range="-0.349066 0.349066"/>
        <position kp="1" ctrlrange="-0.349066 0.349066"/>
      </default>
      <default class="proximal">
        <joint range="-0.261799 1.5708"/>
        <position kp="1" ctrlrange="-0.261799 1.5708"/>
      </default>
      <default class="middle_distal">
        <joint range="0 1.5708"/>
        <position kp="1" ctrlrange="0 3.1415"/>
      </default>
      <default class="plastic">
        <geom solimp="0.5 0.99 0.0001" solref="0.005 1"/>
        <default class="plastic_visual">
          <geom type="mesh" material="black" contype="0" conaffinity="0" group="2"/>
        </default>
        <default class="plastic_collision">
          <geom group="3"/>
        </default>
      </default>
    </default>
  </default>
  <asset>
    <material name="black" specular="0.5" shininess="0.25" rgba="0.16355 0.16355 0.16355 1"/>
    <mesh class="right_hand" file="forearm_collision.obj"/>
    <mesh class="right_hand" file="palm.obj"/>
    <mesh class="right_hand" file="f_proximal.obj"/>
    <mesh class="right_hand" file="f_distal_pst.obj"/>
  </asset>
  <contact>
    <exclude body1="rh_wrist" body2="rh_forearm"/>
  </contact>
  <worldbody>
    <body name="rh_forearm" childclass="right_hand" quat="0 1 0 1" pos="0 0 0.3">
      <inertial mass="3" pos="0 0 0.09" diaginertia="0.0138 0.0138 0.00744"/>
      <geom class="plastic_visual" mesh="forearm_collision"/>
      <geom class="plastic_collision" type="mesh" mesh="forearm_collision"/>
      <body name="rh_wrist" pos="0.01 0 0.21301" quat="1 0 0 1">
        <inertial mass="0.1" pos="0 0 0.029" quat="0.5 0.5 0.5 0.5" diaginertia="6.4e-05 4.38e-05 3.5e-05"/>
        <joint class="wrist_y" name="rh_WRJ2"/>
        <geom class="plastic_collision" size="0.0135 0.015" quat="1 1 0 0" type="capsule"/>
        <body name="rh_palm" pos="0 0 0.034">
          <inertial mass="0.3" pos="0 0 0.035" quat="1 0 0 1" diaginertia="0.0005287 0.0003581 0.000191"/>
          <joint class="wrist_x" name="rh_WRJ1"/>
          <site name="grasp_site" pos="0 -.035 0.09" group="4"/>
          <geom class="plastic_visual" mesh="palm"/>
          <geom class="plastic_collision" size="0.031 0.0035 0.049" pos="0.011 0.0085 0.038" type="box"/>
          <body name="rh_ffknuckle" pos="0.033 0 0.095">
            <inertial mass="0.008" pos="0 0 0" quat="0.5 0.5 -0.5 0.5" diaginertia="3.2e-07 2.6e-07 2.6e-07"/>
            <joint class="knuckle" name="rh_FFJ4"/>
            <body name="rh_ffproximal">
              <inertial mass="0.03" pos="0 0 0.0225" quat="1 0 0 1" diaginertia="1e-05 9.8e-06 1.8e-06"/>
              <joint class="proximal" name="rh_FFJ3"/>
              <geom class="plastic_visual" mesh="f_proximal"/>
              <geom class="plastic_collision" size="0.009 0.02" pos="0 0 0.025" type="capsule"/>
              <body name="rh_ffmiddle" pos="0 0 0.045">
                <inertial mass="0.017" pos="0 0 0.0125" quat="1 0 0 1" diaginertia="2.7e-06 2.6e-06 8.7e-07"/>
                <joint class="middle_distal" name="rh_FFJ2"/>
                <geom class="plastic_collision" size="0.009 0.0125" pos="0 0 0.0125" type="capsule"/>
                <body name="rh_ffdistal" pos="0 0 0.025">
                  <inertial mass="0.013" pos="0 0 0.0130769" quat="1 0 0 1" diaginertia="1.28092e-06 1.12092e-06 5.3e-07"/>
                  <joint class="middle_distal" name="rh_FFJ1"/>
                  <geom class="plastic_visual" mesh="f_distal_pst"/>
                  <geom class="plastic_collision" type="mesh" mesh="f_distal_pst"/>
                </body>
              </body>
            </body>
          </body>
        </body>
      </body>
    </body>
    <geom name="floor" type="plane" size="0 0 0.05"/>
  </worldbody>
  <tendon>
    <fixed name="rh_FFT1">
      <joint joint="rh_FFJ2" coef="1"/>
      <joint joint="rh_FFJ1" coef="1"/>
    </fixed>
  </tendon>
  <actuator>
    <position name="rh_A_WRJ2" joint="rh_WRJ2" class="wrist_y"/>
    <position name="rh_A_WRJ1" joint="rh_WRJ1" class="wrist_x"/>
    <position name="rh_A_FFJ4" joint="rh_FFJ4" class="knuckle"/>
    <position name="rh_A_FFJ3" joint="rh_FFJ3" class="proximal"/>
    <position name="rh_A_FFJ0" tendon="rh_FFT1" class="middle_distal"/>
  </actuator>
  <sensor>
    <jointpos name="rh_FFJ3_pos" joint="rh_FFJ3"/>
    <actuatorfrc name="rh_A_FFJ0_frc" actuator="rh_A_FFJ0"/>
  </sensor>
</mujoco>
'''


def _write_hand_assets(d):
    (d / "assets").mkdir(exist_ok=True)
    faces = [(0, 1, 3), (0, 3, 2), (4, 6, 7), (4, 7, 5), (0, 4, 5), (0, 5, 1), (2, 3, 7), (2, 7, 6), (0, 2, 6), (0, 6, 4), (1, 5, 7),
             (1, 7, 3)]
    for n, h in [("forearm_collision", (0.03, 0.03, 0.08)), ("palm", (0.04, 0.01, 0.05)), ("f_proximal", (0.008, 0.008, 0.02)),
                 ("f_distal_pst", (0.007, 0.007, 0.012))]:
        c = [(sx * h[0], sy * h[1], sz * h[2]) for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
        (d / "assets" / f"{n}.obj").write_text("# v\n" + "".join(f"v {v[0]} {v[1]} {v[2]}\n" for v in c) +
                                              "".join(f"f {a + 1} {b + 1} {e + 1}\n" for a, b, e in faces))


@pytest.fixture()
def hand(capi, tmp_path):
    _write_hand_assets(tmp_path)
    p = tmp_path / "hand.xml"
    p.write_text(HAND_XML)
    return capi.Model.from_xml_file(str(p))


def test_menagerie_style_shadow_hand_compiles_verbatim(hand):
    m = hand
    assert (m.nq, m.nv, m.nu, m.nbody, m.ngeom, m.ntendon, m.nexclude, m.nsensor) == (6, 6, 5, 8, 11, 1, 1, 2)
    assert (m.opt.impratio, m.opt.cone) == (10.0, 1)
    # class chain right_hand -> wrist -> wrist_y / wrist_x; right_hand -> knuckle / proximal / middle_distal
    np.testing.assert_allclose(m.dof_damping, [0.5, 0.5, 0.05, 0.05, 0.05, 0.05])
    np.testing.assert_allclose(m.dof_frictionloss, 0.01)
    np.testing.assert_allclose(m.dof_armature, 0.0002)
    np.testing.assert_allclose(m.jnt_axis, [[0, 1, 0], [1, 0, 0], [0, -1, 0], [1, 0, 0], [1, 0, 0], [1, 0, 0]])
    np.testing.assert_allclose(m.jnt_range[:, 1], [0.174533, 0.488692, 0.349066, 1.5708, 1.5708, 1.5708])
    # <position class=...>: kp as gain and -kp as the position bias, limits from the class, the tendon-driven one too
    np.testing.assert_allclose(m.actuator_gainprm[:, 0], [10, 8, 1, 1, 1])
    np.testing.assert_allclose(m.actuator_biasprm[:, 1], [-10, -8, -1, -1, -1])
    np.testing.assert_allclose(m.actuator_forcerange, [[-10, 10], [-5, 5], [-1, 1], [-1, 1], [-1, 1]])
    np.testing.assert_allclose(m.actuator_ctrlrange[4], [0, 3.1415])
    assert m.actuator_trntype.tolist() == [0, 0, 0, 0, 3]
    # plastic class on every hand geom (visual ones do not collide); the floor keeps the global defaults
    assert m.geom_contype.tolist() == [1, 0, 1, 1, 0, 1, 0, 1, 1, 0, 1]
    np.testing.assert_allclose(m.geom_solref[0], [0.02, 1])
    np.testing.assert_allclose(m.geom_solref[1:], np.tile([0.005, 1], (10, 1)))
    np.testing.assert_allclose(m.geom_solimp[1:, :3], np.tile([0.5, 0.99, 0.0001], (10, 1)))


@pytest.mark.gpu
def test_menagerie_style_shadow_hand_gpu_parity(hand, capi, orc):
    from mujoco_ros_pkgs_b200.batch import BatchSim
    from parity_util import compare_forward_fields, injected_steps, make_oracles, perturbed

    model, nenv = hand, 6
    qpos, qvel = perturbed(model, nenv, seed=8, amp=0.1)
    qpos = np.clip(qpos, model.jnt_range[:, 0], model.jnt_range[:, 1])
    sim = BatchSim(model, nenv)
    sim.set("qpos", qpos)
    sim.set("qvel", qvel)
    sim.keep_intermediates(True)
    sim.forward()
    oracles = make_oracles(orc, model, qpos, qvel)
    for o in oracles:
        o.forward()
    compare_forward_fields(capi, model, sim, oracles, skip={"xfrc_applied"}, tag="menagerie-hand")
    sim.keep_intermediates(False)
    worst, _ = injected_steps(model, sim, oracles, 150, np.random.default_rng(4), tag="menagerie-hand")
    assert worst < 1e-5
