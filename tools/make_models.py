"""Authors the BASELINE config models the reference does not ship: a Shadow-Hand-like
24-DoF hand (C3), a 27-DoF humanoid with the C4 sensor suite, and the 20-box cluttered bin (C5).
Primitive colliders only.  Output: mujoco_ros_pkgs_b200/models/{hand_like,humanoid_like,bin}.xml"""
import os

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mujoco_ros_pkgs_b200", "models")


def hand():
    L = []
    A = L.append
    A('<!-- Shadow-Hand-like hand (BASELINE config C3): 24 DoF (wrist 2, FF/MF/RF 4 each, LF 5, thumb 5), 20 actuators,')
    A('     4 fixed tendons coupling the two distal joints of the four fingers, capsule links, RK4 + Newton + elliptic. -->')
    A('<mujoco model="hand_like">')
    A('  <compiler angle="radian" autolimits="true"/>')
    A('  <size nconmax="32" njmax="96"/>')
    A('  <option timestep="0.002" integrator="RK4" solver="Newton" cone="elliptic" gravity="0 0 -9.81"/>')
    A('  <default>')
    A('    <joint type="hinge" damping="0.05" armature="0.0002" limited="true"/>')
    A('    <geom type="capsule" size="0.009" friction="1 0.005 0.0001" density="800"/>')
    A('    <motor ctrllimited="true" ctrlrange="-1 1"/>')
    A('  </default>')
    A('  <worldbody>')
    A('    <geom name="floor" type="plane" size="1 1 0.1" pos="0 0 -0.25"/>')
    A('    <body name="forearm" pos="0 0 0.1">')
    A('      <geom name="forearm_c" fromto="0 0 -0.1 0 0 0" size="0.03" contype="0" conaffinity="0"/>')
    A('      <body name="wrist" pos="0 0 0.01">')
    A('        <joint name="WRJ1" axis="0 1 0" range="-0.49 0.14" damping="0.5" armature="0.005"/>')
    A('        <geom name="wrist_c" type="sphere" size="0.02" contype="0" conaffinity="0"/>')
    A('        <body name="palm" pos="0 0 0.034">')
    A('          <joint name="WRJ0" axis="1 0 0" range="-0.7 0.49" damping="0.5" armature="0.005"/>')
    A('          <geom name="palm_c" type="box" size="0.04 0.012 0.045" pos="0 0 0.04"/>')
    fingers = [("FF", 0.033, 0.095, False), ("MF", 0.011, 0.099, False), ("RF", -0.011, 0.095, False), ("LF", -0.033, 0.0866, True)]
    for name, x, z, meta in fingers:
        ind = "          "
        if meta:
            A(f'{ind}<body name="{name}meta" pos="{x} 0 0.02">')
            A(f'{ind}  <joint name="{name}J4" axis="0.57 0 0.82" range="0 0.785"/>')
            A(f'{ind}  <geom name="{name}meta_c" fromto="0 0 0 0 0 {z - 0.02:.4f}" size="0.009" contype="0" conaffinity="0"/>')
            A(f'{ind}  <body name="{name}knuckle" pos="0 0 {z - 0.02:.4f}">')
            ind2 = ind + "    "
        else:
            A(f'{ind}<body name="{name}knuckle" pos="{x} 0 {z}">')
            ind2 = ind + "  "
        A(f'{ind2}<joint name="{name}J3" axis="0 1 0" range="-0.349 0.349"/>')
        A(f'{ind2}<geom name="{name}knuckle_c" type="sphere" size="0.0095" contype="0" conaffinity="0"/>')
        A(f'{ind2}<body name="{name}proximal">')
        A(f'{ind2}  <joint name="{name}J2" axis="1 0 0" range="0 1.571"/>')
        A(f'{ind2}  <geom name="{name}proximal_c" fromto="0 0 0 0 0 0.045"/>')
        A(f'{ind2}  <body name="{name}middle" pos="0 0 0.045">')
        A(f'{ind2}    <joint name="{name}J1" axis="1 0 0" range="0 1.571"/>')
        A(f'{ind2}    <geom name="{name}middle_c" fromto="0 0 0 0 0 0.025" size="0.0085"/>')
        A(f'{ind2}    <body name="{name}distal" pos="0 0 0.025">')
        A(f'{ind2}      <joint name="{name}J0" axis="1 0 0" range="0 1.571"/>')
        A(f'{ind2}      <geom name="{name}distal_c" fromto="0 0 0 0 0 0.024" size="0.008"/>')
        A(f'{ind2}      <site name="{name}tip" pos="0 0 0.026" size="0.004"/>')
        A(f'{ind2}    </body>')
        A(f'{ind2}  </body>')
        A(f'{ind2}</body>')
        A(f'{ind}{"  " if meta else ""}</body>')
        if meta:
            A(f'{ind}</body>')
    ind = "          "
    A(f'{ind}<body name="THbase" pos="0.034 -0.009 0.029" quat="0.9239 0 0.3827 0">')
    A(f'{ind}  <joint name="THJ4" axis="0 0 -1" range="-1.047 1.047"/>')
    A(f'{ind}  <geom name="THbase_c" type="sphere" size="0.011" contype="0" conaffinity="0"/>')
    A(f'{ind}  <body name="THproximal">')
    A(f'{ind}    <joint name="THJ3" axis="1 0 0" range="0 1.222"/>')
    A(f'{ind}    <geom name="THproximal_c" fromto="0 0 0 0 0 0.038" size="0.011"/>')
    A(f'{ind}    <body name="THhub" pos="0 0 0.038">')
    A(f'{ind}      <joint name="THJ2" axis="1 0 0" range="-0.209 0.209"/>')
    A(f'{ind}      <geom name="THhub_c" type="sphere" size="0.0105" contype="0" conaffinity="0"/>')
    A(f'{ind}      <body name="THmiddle">')
    A(f'{ind}        <joint name="THJ1" axis="0 1 0" range="-0.524 0.524"/>')
    A(f'{ind}        <geom name="THmiddle_c" fromto="0 0 0 0 0 0.032" size="0.01"/>')
    A(f'{ind}        <body name="THdistal" pos="0 0 0.032">')
    A(f'{ind}          <joint name="THJ0" axis="0 1 0" range="-1.571 0"/>')
    A(f'{ind}          <geom name="THdistal_c" fromto="0 0 0 0 0 0.026" size="0.009"/>')
    A(f'{ind}          <site name="THtip" pos="0 0 0.028" size="0.004"/>')
    A(f'{ind}        </body>')
    A(f'{ind}      </body>')
    A(f'{ind}    </body>')
    A(f'{ind}  </body>')
    A(f'{ind}</body>')
    A('        </body>')
    A('      </body>')
    A('    </body>')
    A('  </worldbody>')
    A('  <contact>')
    A('    <exclude body1="THproximal" body2="THmiddle"/>  <!-- joined through the collision-free hub: coincident segment ends -->')
    A('  </contact>')
    A('  <tendon>')
    for name, _, _, _ in fingers:
        A(f'    <fixed name="{name}T"><joint joint="{name}J1" coef="1"/><joint joint="{name}J0" coef="1"/></fixed>')
    A('  </tendon>')
    A('  <actuator>')
    A('    <motor name="A_WRJ1" joint="WRJ1" gear="5"/>')
    A('    <motor name="A_WRJ0" joint="WRJ0" gear="5"/>')
    for name, _, _, meta in fingers:
        if meta:
            A(f'    <motor name="A_{name}J4" joint="{name}J4" gear="1"/>')
        A(f'    <motor name="A_{name}J3" joint="{name}J3" gear="1"/>')
        A(f'    <motor name="A_{name}J2" joint="{name}J2" gear="1"/>')
        A(f'    <motor name="A_{name}T" tendon="{name}T" gear="0.7"/>')
    for k in (4, 3, 2, 1, 0):
        A(f'    <motor name="A_THJ{k}" joint="THJ{k}" gear="1"/>')
    A('  </actuator>')
    A('</mujoco>')
    return "\n".join(L) + "\n"


def humanoid():
    return '''<!-- Humanoid (BASELINE config C4): 17 bodies, free root + 21 hinges = 27 DoF, 21 motors, capsule / sphere
     colliders on a plane, and the C4 sensor suite (accelerometer, gyro, velocimeter, framepos, framequat,
     subtreecom, 21 jointpos, 21 jointvel, 2 touch; nsensordata = 63).  Euler, Newton, pyramidal. -->
<mujoco model="humanoid_like">
  <compiler angle="degree" autolimits="true"/>
  <size nconmax="32" njmax="128"/>
  <option timestep="0.005" integrator="Euler" solver="Newton" cone="pyramidal"/>
  <default>
    <joint type="hinge" damping="0.2" stiffness="1" armature="0.01" limited="true" solimplimit="0 0.99 0.01"/>
    <geom type="capsule" condim="3" friction="0.7 0.005 0.0001" solref="0.015 1" solimp="0.99 0.99 0.003"/>
    <motor ctrllimited="true" ctrlrange="-1 1"/>
  </default>
  <worldbody>
    <geom name="floor" type="plane" size="10 10 0.1" condim="3"/>
    <body name="torso" pos="0 0 1.282">
      <freejoint name="root"/>
      <site name="imu" pos="0 0 0" size="0.01"/>
      <geom name="torso" fromto="0 -0.07 0 0 0.07 0" size="0.07"/>
      <geom name="upper_waist" fromto="-0.01 -0.06 -0.12 -0.01 0.06 -0.12" size="0.06"/>
      <body name="head" pos="0 0 0.19">
        <geom name="head" type="sphere" size="0.09"/>
      </body>
      <body name="lower_waist" pos="-0.01 0 -0.26" quat="1 0 -0.002 0">
        <joint name="abdomen_z" pos="0 0 0.065" axis="0 0 1" range="-45 45" damping="5" stiffness="20" armature="0.02"/>
        <joint name="abdomen_y" pos="0 0 0.065" axis="0 1 0" range="-75 30" damping="5" stiffness="10" armature="0.02"/>
        <geom name="lower_waist" fromto="0 -0.06 0 0 0.06 0" size="0.06"/>
        <body name="pelvis" pos="0 0 -0.165" quat="1 0 -0.002 0">
          <joint name="abdomen_x" pos="0 0 0.1" axis="1 0 0" range="-35 35" damping="5" stiffness="10" armature="0.02"/>
          <geom name="butt" fromto="-0.02 -0.07 0 -0.02 0.07 0" size="0.09"/>
          <body name="right_thigh" pos="0 -0.1 -0.04">
            <joint name="right_hip_x" axis="1 0 0" range="-25 5" damping="5" stiffness="10" armature="0.01"/>
            <joint name="right_hip_z" axis="0 0 1" range="-60 35" damping="5" stiffness="10" armature="0.01"/>
            <joint name="right_hip_y" axis="0 1 0" range="-110 20" damping="5" stiffness="20" armature="0.008"/>
            <geom name="right_thigh" fromto="0 0 0 0 0.01 -0.34" size="0.06"/>
            <body name="right_shin" pos="0 0.01 -0.403">
              <joint name="right_knee" pos="0 0 0.02" axis="0 -1 0" range="-160 -2" armature="0.006"/>
              <geom name="right_shin" fromto="0 0 0 0 0 -0.3" size="0.049"/>
              <body name="right_foot" pos="0 0 -0.39">
                <joint name="right_ankle_y" pos="0 0 0.08" axis="0 1 0" range="-50 50" stiffness="4" armature="0.0008"/>
                <joint name="right_ankle_x" pos="0 0 0.04" axis="1 0 0.5" range="-50 50" stiffness="1" armature="0.0006"/>
                <geom name="right_foot_a" fromto="-0.07 -0.02 0 0.14 -0.04 0" size="0.027"/>
                <geom name="right_foot_b" fromto="-0.07 0 0 0.14 0.02 0" size="0.027"/>
                <site name="right_sole" pos="0.035 -0.01 0" size="0.12 0.06 0.04" type="box"/>
              </body>
            </body>
          </body>
          <body name="left_thigh" pos="0 0.1 -0.04">
            <joint name="left_hip_x" axis="-1 0 0" range="-25 5" damping="5" stiffness="10" armature="0.01"/>
            <joint name="left_hip_z" axis="0 0 -1" range="-60 35" damping="5" stiffness="10" armature="0.01"/>
            <joint name="left_hip_y" axis="0 1 0" range="-110 20" damping="5" stiffness="20" armature="0.008"/>
            <geom name="left_thigh" fromto="0 0 0 0 -0.01 -0.34" size="0.06"/>
            <body name="left_shin" pos="0 -0.01 -0.403">
              <joint name="left_knee" pos="0 0 0.02" axis="0 -1 0" range="-160 -2" armature="0.006"/>
              <geom name="left_shin" fromto="0 0 0 0 0 -0.3" size="0.049"/>
              <body name="left_foot" pos="0 0 -0.39">
                <joint name="left_ankle_y" pos="0 0 0.08" axis="0 1 0" range="-50 50" stiffness="4" armature="0.0008"/>
                <joint name="left_ankle_x" pos="0 0 0.04" axis="1 0 0.5" range="-50 50" stiffness="1" armature="0.0006"/>
                <geom name="left_foot_a" fromto="-0.07 0.02 0 0.14 0.04 0" size="0.027"/>
                <geom name="left_foot_b" fromto="-0.07 0 0 0.14 -0.02 0" size="0.027"/>
                <site name="left_sole" pos="0.035 0.01 0" size="0.12 0.06 0.04" type="box"/>
              </body>
            </body>
          </body>
        </body>
      </body>
      <body name="right_upper_arm" pos="0 -0.17 0.06">
        <joint name="right_shoulder1" axis="2 1 1" range="-85 60" stiffness="1" armature="0.0068"/>
        <joint name="right_shoulder2" axis="0 -1 1" range="-85 60" stiffness="1" armature="0.0051"/>
        <geom name="right_uarm" fromto="0 0 0 0.16 -0.16 -0.16" size="0.04"/>
        <body name="right_lower_arm" pos="0.18 -0.18 -0.18">
          <joint name="right_elbow" axis="0 -1 1" range="-90 50" stiffness="0" armature="0.0028"/>
          <geom name="right_larm" fromto="0.01 0.01 0.01 0.17 0.17 0.17" size="0.031"/>
          <body name="right_hand" pos="0.18 0.18 0.18">
            <geom name="right_hand" type="sphere" size="0.04"/>
          </body>
        </body>
      </body>
      <body name="left_upper_arm" pos="0 0.17 0.06">
        <joint name="left_shoulder1" axis="2 -1 1" range="-60 85" stiffness="1" armature="0.0068"/>
        <joint name="left_shoulder2" axis="0 1 1" range="-60 85" stiffness="1" armature="0.0051"/>
        <geom name="left_uarm" fromto="0 0 0 0.16 0.16 -0.16" size="0.04"/>
        <body name="left_lower_arm" pos="0.18 0.18 -0.18">
          <joint name="left_elbow" axis="0 -1 -1" range="-90 50" stiffness="0" armature="0.0028"/>
          <geom name="left_larm" fromto="0.01 -0.01 0.01 0.17 -0.17 0.17" size="0.031"/>
          <body name="left_hand" pos="0.18 -0.18 0.18">
            <geom name="left_hand" type="sphere" size="0.04"/>
          </body>
        </body>
      </body>
    </body>
  </worldbody>
  <actuator>
''' + "".join(f'    <motor name="{j}" joint="{j}" gear="{g}"/>\n' for j, g in [
        ("abdomen_y", 100), ("abdomen_z", 100), ("abdomen_x", 100), ("right_hip_x", 100), ("right_hip_z", 100),
        ("right_hip_y", 300), ("right_knee", 200), ("right_ankle_y", 50), ("right_ankle_x", 50), ("left_hip_x", 100),
        ("left_hip_z", 100), ("left_hip_y", 300), ("left_knee", 200), ("left_ankle_y", 50), ("left_ankle_x", 50),
        ("right_shoulder1", 25), ("right_shoulder2", 25), ("right_elbow", 25), ("left_shoulder1", 25),
        ("left_shoulder2", 25), ("left_elbow", 25)]) + '''  </actuator>
  <sensor>
    <accelerometer name="torso_acc" site="imu"/>
    <gyro name="torso_gyro" site="imu"/>
    <velocimeter name="torso_vel" site="imu"/>
    <framepos name="torso_pos" objtype="site" objname="imu"/>
    <framequat name="torso_quat" objtype="site" objname="imu"/>
    <subtreecom name="com" body="torso"/>
''' + "".join(f'    <jointpos name="jp_{j}" joint="{j}"/>\n' for j in JOINTS) + "".join(
        f'    <jointvel name="jv_{j}" joint="{j}"/>\n' for j in JOINTS) + '''    <touch name="right_touch" site="right_sole"/>
    <touch name="left_touch" site="left_sole"/>
  </sensor>
</mujoco>
'''


JOINTS = ["abdomen_z", "abdomen_y", "abdomen_x", "right_hip_x", "right_hip_z", "right_hip_y", "right_knee",
          "right_ankle_y", "right_ankle_x", "left_hip_x", "left_hip_z", "left_hip_y", "left_knee", "left_ankle_y",
          "left_ankle_x", "right_shoulder1", "right_shoulder2", "right_elbow", "left_shoulder1", "left_shoulder2",
          "left_elbow"]


def bin_model():
    rng = np.random.default_rng(77)
    L = []
    A = L.append
    A('<!-- Cluttered bin (BASELINE config C5): plane + 4 static wall boxes + 20 free boxes (half-extents U(0.02,0.04))')
    A('     dropped from a 4 x 5 grid at z in [0.1, 0.5] with random orientation.  Euler, Newton, elliptic. -->')
    A('<mujoco model="bin">')
    A('  <size nconmax="160" njmax="480"/>')
    A('  <option timestep="0.002" integrator="Euler" solver="Newton" cone="elliptic" iterations="100"/>')
    A('  <default><geom type="box" friction="0.8 0.005 0.0001"/></default>')
    A('  <worldbody>')
    A('    <geom name="floor" type="plane" size="1 1 0.1"/>')
    A('    <geom name="wall_px" size="0.01 0.26 0.15" pos="0.25 0 0.15"/>')
    A('    <geom name="wall_nx" size="0.01 0.26 0.15" pos="-0.25 0 0.15"/>')
    A('    <geom name="wall_py" size="0.26 0.01 0.15" pos="0 0.25 0.15"/>')
    A('    <geom name="wall_ny" size="0.26 0.01 0.15" pos="0 -0.25 0.15"/>')
    k = 0
    for ix in range(4):
        for iy in range(5):
            s = rng.uniform(0.02, 0.04, 3)
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            x, y = -0.15 + 0.1 * ix, -0.16 + 0.08 * iy
            z = rng.uniform(0.1, 0.5)
            A(f'    <body name="box{k}" pos="{x:.4f} {y:.4f} {z:.4f}" quat="{q[0]:.5f} {q[1]:.5f} {q[2]:.5f} {q[3]:.5f}">')
            A(f'      <freejoint name="box{k}_j"/>')
            A(f'      <geom name="box{k}_g" size="{s[0]:.4f} {s[1]:.4f} {s[2]:.4f}"/>')
            A('    </body>')
            k += 1
    A('  </worldbody>')
    A('</mujoco>')
    return "\n".join(L) + "\n"


if __name__ == "__main__":
    for name, text in (("hand_like.xml", hand()), ("humanoid_like.xml", humanoid()), ("bin.xml", bin_model())):
        with open(os.path.join(OUT, name), "w") as f:
            f.write(text)
        print("wrote", name)
