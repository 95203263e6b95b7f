import sys, torch
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from helpers import build_denoiser
from rald_b200 import synth
net = build_denoiser(device='cuda:0')
cube = synth.radar_cube(2, seed=3).cuda()
outs = [net.process_radar_cond(cube).clone() for _ in range(4)]
for o in outs[1:]:
    print('tokens equal', torch.equal(o, outs[0]), float((o - outs[0]).abs().max()))
cube2 = cube.clone()
o2 = net.process_radar_cond(cube2)
print('clone input equal', torch.equal(o2, outs[0]), float((o2 - outs[0]).abs().max()))
enc = [net.radar_enc(cube[..., 0:1].permute(0, 4, 1, 2, 3)).clone() for _ in range(3)] if hasattr(net, 'radar_enc') else []
for e in enc[1:]:
    print('enc equal', torch.equal(e, enc[0]), float((e - enc[0]).abs().max()))
