import os, sys, numpy as np
os.environ["EVG_STEP_KERNEL"] = "pair"
sys.path.insert(0, "/root/repo")
import torch, evgsim
env = evgsim.BatchedEvergladesEnv(int(sys.argv[1]) if len(sys.argv) > 1 else 16, seed=1)
env.reset()
torch.cuda.synchronize()
print("reset ok", flush=True)
for t in range(int(sys.argv[2]) if len(sys.argv) > 2 else 3):
    env.step(env.random_actions())
    torch.cuda.synchronize()
    print("step", t, "ok", flush=True)
