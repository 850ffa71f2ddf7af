import sys,os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
import numpy as np
from mujoco_ros_pkgs_b200 import _capi
from mujoco_ros_pkgs_b200.batch import BatchSim
from oracle import binding as ob
import gpu_check
m=_capi.Model.from_xml_file('/root/repo/mujoco_ros_pkgs_b200/models/hand_like.xml')
nenv=8
qpos,qvel=gpu_check.perturb(m,nenv,0.02)
sim=BatchSim(m,nenv); sim.set('qpos',qpos); sim.set('qvel',qvel); sim.keep_intermediates(True); sim.forward()
for e in range(nenv):
    o=ob.Oracle(m); o.set('qpos',qpos[e]); o.set('qvel',qvel[e]); o.forward()
    nc=o.get('ncon')[0]
    gf=sim.get('contact_frame')[e][:9*nc].reshape(nc,9); of=o.get('contact_frame')[:9*nc].reshape(nc,9)
    gp=sim.get('contact_pos')[e][:3*nc].reshape(nc,3); op=o.get('contact_pos')[:3*nc].reshape(nc,3)
    gd=sim.get('contact_dist')[e][:nc]; od=o.get('contact_dist')[:nc]
    g1=o.get('contact_geom1')[:nc]; g2=o.get('contact_geom2')[:nc]
    for c in range(nc):
        d=np.abs(gf[c]-of[c]).max()
        if d>1e-9 or np.abs(gp[c]-op[c]).max()>1e-9:
            print('env',e,'con',c,'geoms',m.id2name(5,int(g1[c])),m.geom_type[g1[c]],m.id2name(5,int(g2[c])),m.geom_type[g2[c]],'dist',gd[c],od[c])
            print('   gpu frame',gf[c][:6],'pos',gp[c]); print('   orc frame',of[c][:6],'pos',op[c])
