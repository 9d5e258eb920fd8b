import sys, os, json
sys.path.insert(0, "/root/repo")
import torch, evgsim
from evgsim import policy
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
env = evgsim.BatchedEvergladesEnv(n, seed=0, auto_reset=1)
env.reset()
for _ in range(40): env.step(env.random_actions())
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(105, 528), torch.nn.ReLU(), torch.nn.Linear(528, 132)).cuda()
f = policy.FusedDQN(env, net)
def t(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
x = env.obs.view(-1, 105)
nb = net.to(torch.bfloat16)
print(json.dumps({"rows": 2 * n, "fused_mlp_us": t(f.forward), "fused_mlp_t_us": t(f.forward_t), "decode_us": t(lambda: env.decode_dqn(f.q)), "mlp_t_plus_decode_us": t(f),
                  "torch_bf16_us": t(lambda: nb(x.to(torch.bfloat16))), "step_us": t(lambda: env.step(env._actions)),
                  "tflops_fused": 2 * 2 * n * (128 * 576 + 576 * 144) / (t(f.forward) * 1e-6) / 1e12}))
