"""Writes the seed corpus for tools/fuzz_readers.cpp into a directory: the shipped models, one scene that names a mesh of
each format (binary STL, ASCII STL, OBJ), a PNG and a binary height field, keyframes / tendons / sensors / defaults, and
the binary model file of one of them."""
import os
import shutil
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main(out):
    from test_hfield import _png
    from mujoco_ros_pkgs_b200 import _capi as capi
    os.makedirs(out, exist_ok=True)
    for f in ("pendulum_scene.xml", "equality_scene.xml", "actuated_arm.xml", "panda_like.xml"):
        shutil.copy(os.path.join(ROOT, "mujoco_ros_pkgs_b200", "models", f), out)
    rng = np.random.default_rng(0)
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float) * 0.1
    faces = [(0, 2, 1), (0, 1, 3), (0, 3, 2), (1, 2, 3)]
    with open(os.path.join(out, "tet.stl"), "wb") as fp:
        fp.write(b"\0" * 80 + struct.pack("<I", len(faces)))
        for f in faces:
            fp.write(struct.pack("<3f", 0, 0, 0) + b"".join(struct.pack("<3f", *v[i]) for i in f) + b"\0\0")
    with open(os.path.join(out, "tet_ascii.stl"), "w") as fp:
        fp.write("solid t\n")
        for f in faces:
            fp.write("facet normal 0 0 0\nouter loop\n" + "".join("vertex %g %g %g\n" % tuple(v[i]) for i in f) + "endloop\nendfacet\n")
        fp.write("endsolid t\n")
    with open(os.path.join(out, "tet.obj"), "w") as fp:
        fp.write("".join("v %g %g %g\n" % tuple(p) for p in v) + "".join("f %d/1/1 %d %d\n" % tuple(i + 1 for i in f) for f in faces))
    open(os.path.join(out, "hill.png"), "wb").write(_png(rng.integers(0, 256, (9, 8, 1)), 0, 8))
    with open(os.path.join(out, "hill.bin"), "wb") as fp:
        fp.write(struct.pack("<2i", 4, 5) + rng.random(20).astype("<f4").tobytes())
    xml = """<mujoco model="seed">
  <compiler angle="degree" autolimits="true"/>
  <option timestep="0.002" cone="elliptic" solver="Newton"><flag energy="enable"/></option>
  <default><default class="a"><geom friction="0.8 0.01 0.001" condim="4"/><joint damping="0.2"/></default></default>
  <asset>
    <mesh name="m1" file="tet.stl" scale="2 2 2"/><mesh name="m2" file="tet_ascii.stl"/><mesh name="m3" file="tet.obj"/>
    <mesh name="m4" vertex="0 0 0 .1 0 0 0 .1 0 0 0 .1 .1 .1 .1"/>
    <hfield name="h1" file="hill.png" size="1 1 .2 .1"/><hfield name="h2" file="hill.bin" size=".5 .5 .1 .1"/>
  </asset>
  <worldbody>
    <geom type="hfield" hfield="h1"/><geom type="hfield" hfield="h2" pos="3 0 0"/>
    <body name="b1" pos="0 0 1" childclass="a"><freejoint/><geom type="mesh" mesh="m1"/><site name="s1" pos="0 0 .1"/>
      <body name="b2" pos=".2 0 0"><joint name="j1" axis="0 1 0" range="-45 45" springdamper="0.1 1"/><geom type="mesh" mesh="m2"/>
        <geom type="capsule" fromto="0 0 0 .1 0 0" size=".02"/><site name="s2" pos=".1 0 0"/></body></body>
    <body name="b3" pos="1 0 1"><joint name="j2" type="ball"/><geom type="mesh" mesh="m3"/><geom type="ellipsoid" size=".1 .05 .03"/></body>
    <body name="b4" pos="0 1 1"><joint name="j3" type="slide" axis="0 0 1"/><geom type="mesh" mesh="m4"/><geom type="cylinder" size=".05 .1"/></body>
    <frame pos="0 0 .5" euler="0 0 45" childclass="a"><geom size=".02" pos="2 0 0"/><frame pos=".1 0 0"><body name="fb" pos="0 2 0" gravcomp="1">
      <joint name="jf" axis="1 0 0"/><geom size=".05"/></body></frame></frame>
    <body name="mc" mocap="true" pos="0 0 2"><geom type="box" size=".05 .05 .05" contype="0" conaffinity="0"/></body>
  </worldbody>
  <contact><exclude body1="b1" body2="b2"/><pair geom1="1" geom2="0" condim="3"/></contact>
  <tendon><spatial name="t1" stiffness="10"><site site="s1"/><site site="s2"/></spatial>
    <fixed name="t2"><joint joint="j1" coef="1"/><joint joint="j3" coef="-2"/></fixed></tendon>
  <equality><weld body1="mc" body2="b3"/><joint joint1="j1" joint2="j3" polycoef="0 1 0 0 0"/></equality>
  <actuator><position joint="j1" kp="5"/><intvelocity joint="j3" kp="3" actrange="-1 1"/><damper joint="j1" kv="1" ctrlrange="0 1"/>
    <cylinder tendon="t2" timeconst=".1" area=".01"/><general tendon="t1" gaintype="affine" gainprm="1 .1 .1"/></actuator>
  <sensor><jointpos joint="j1"/><framepos objtype="site" objname="s1"/><accelerometer site="s2"/><tendonpos tendon="t1"/>
    <actuatorfrc actuator="0"/></sensor>
  <keyframe><key name="k" time="1" qpos="0 0 1 1 0 0 0 .1 1 0 0 0 .2 0" ctrl="1 0 0 0 0"/></keyframe>
</mujoco>
""".replace('geom1="1" geom2="0"', 'geom1="g_a" geom2="g_b"').replace('<geom type="hfield" hfield="h1"/>', '<geom name="g_b" type="hfield" hfield="h1"/>').replace(
        '<geom type="mesh" mesh="m1"/>', '<geom name="g_a" type="mesh" mesh="m1"/>').replace('actuator="0"', 'actuator="a0"').replace(
        '<position joint="j1" kp="5"/>', '<position name="a0" joint="j1" kp="5"/>')
    path = os.path.join(out, "assets_scene.xml")
    open(path, "w").write(xml)
    m = capi.Model.from_xml_file(path)  # the seed itself must load
    m.save_binary(os.path.join(out, "assets_scene.b2mjb"))
    print("seeds in", out, sorted(os.listdir(out)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/tmp/b2mj_fuzz_seeds")
