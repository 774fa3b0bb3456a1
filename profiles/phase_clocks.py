import sys; sys.path.insert(0,'.')
import numpy as np, torch
from mmwave_msc_b200 import synth, pose_weights as pw
from mmwave_msc_b200.batched import BatchedTracker
S=1024; F=30
b = synth.gen_batch(range(S), F)
bt = BatchedTracker(S)
for f in range(20): bt.step(b[f].points, b[f].offsets, b[f].dt, pose=False)
bt.sync(); bt.phase_clocks(True)
for f in range(20,30): bt.step(b[f].points, b[f].offsets, b[f].dt, pose=False)
pc = bt.phase_clocks(False).astype(float)
names=["load+filter","load tracks","predict+gatemat","gate","assoc stats","maintain","update","dbscan","spawn","writeback"]
tot=pc[1:11].sum()
for i,n in enumerate(names): print("%-16s %8.0f cyc/scene-frame %5.1f%%"%(n, pc[i+1]/(S*10), 100*pc[i+1]/tot))
print("total cycles per scene-frame", tot/(S*10))
