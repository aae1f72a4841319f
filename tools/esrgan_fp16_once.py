import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/super-resolution-images-for-3d-printing-defect-detection_b200")
from srb200 import engine, weights
x = torch.rand((512, 24, 24, 3), device="cuda") * 2 - 1
net = engine.ESRGANGeneratorNet(weights.esrgan_generator_weights(2, 8, 4), 2, 8, 4, precision="fp16")
for _ in range(2):
    net.forward_device(x)
torch.cuda.synchronize()
print("ok")
