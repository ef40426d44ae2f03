import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ecologysemanticsegmentation_b200 import test_video as tv
rs = np.random.RandomState(0)
frames = torch.from_numpy((rs.rand(16, 1080, 1920, 3) * 255).astype(np.uint8)).cuda()
for _ in range(3):
    tv.preprocess_frames(frames, (512, 512))
torch.cuda.synchronize()
