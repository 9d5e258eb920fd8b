#!/usr/bin/env python
"""BASELINE.json configs[3]: PyTorch policy-in-the-loop self-play rollout.

    python tools/policy_rollout.py [--envs-per-gpu 32768] [--turns 300] [--policy dqn|ppo|rppo] [--dtype fp32|bf16] [--graph]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/policy_rollout.py

Both players are driven by one network of the reference's shape — DQN: Linear(105, 528) - ReLU - Linear(528, 132)
(agents/DQN/QNetwork.py:37-42, random weights), decoded like DQNAgent.filter_actions (evg_decode_dqn); PPO: the
actor's Linear(105, 128) - Tanh - Linear(128, 128) - Tanh - Linear(128, 132) - Tanh - Softmax head
(agents/PPO/ActorCritic.py:33-50, without the GRU), 7 indices sampled without replacement and unravelled like
PPOAgent.get_action (evg_decode_indices); RPPO: the same head, its output repeated 7 times through the actor's
GRU(128, 128) (hidden state carried over the 7 steps and from turn to turn, zeroed when a match ends), one index
sampled per step (ActorCritic.act with use_recurrent, agents/PPO/ActorCritic.py:79-105).  --graph captures the whole
turn (forward, sampling, decode, step) in a CUDA graph.  The observation tensor the step kernel writes is consumed in place
(a [N*2, 105] view, no copy, no dtype conversion kernel for fp32); the decoded int8 rows go straight back into the
step.  Matches shard over the ranks with no collective on the path; rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import evgsim


class RecurrentActor(torch.nn.Module):
    """ActorCritic's actor with use_recurrent=True (agents/PPO/ActorCritic.py:33-50,79-105), batched over (match, player)."""

    def __init__(self, n_latent=128):
        super().__init__()
        self.action_head = torch.nn.Sequential(torch.nn.Linear(105, n_latent), torch.nn.Tanh(), torch.nn.Linear(n_latent, n_latent), torch.nn.Tanh())
        self.action_gru = torch.nn.GRU(n_latent, n_latent, batch_first=False)
        self.action_layer = torch.nn.Sequential(torch.nn.Linear(n_latent, 132), torch.nn.Tanh(), torch.nn.Softmax(dim=-1))

    def forward(self, x, hidden):
        h = self.action_head(x)                                   # [B, 128]
        seq, hidden = self.action_gru(h.unsqueeze(0).expand(7, -1, -1).contiguous(), hidden)  # the observation repeated for the 7 actions
        return self.action_layer(seq), hidden                     # [7, B, 132], [1, B, 128]


def build_policy(kind, dtype, device, seed=0):
    torch.manual_seed(seed)
    if kind == "dqn":
        net = torch.nn.Sequential(torch.nn.Linear(105, 528), torch.nn.ReLU(), torch.nn.Linear(528, 132))
    elif kind == "rppo":
        net = RecurrentActor()
    else:
        net = torch.nn.Sequential(torch.nn.Linear(105, 128), torch.nn.Tanh(), torch.nn.Linear(128, 128), torch.nn.Tanh(),
                                  torch.nn.Linear(128, 132), torch.nn.Tanh(), torch.nn.Softmax(dim=-1))
    return net.to(device=device, dtype=dtype).eval()


@torch.no_grad()
def policy_actions(env, net, kind, dtype, hidden=None):
    x = env.obs.view(-1, env.obs_len)  # [N*2, 105], the step kernel's output buffer itself
    x = x if dtype == torch.float32 else x.to(dtype)
    if kind == "rppo":
        probs, h = net(x, hidden)
        hidden.copy_(h)
        idx = torch.multinomial(probs.float().view(-1, 132), 1).view(7, env.num_envs, 2).permute(1, 2, 0).contiguous()
        return env.decode_indices(idx, div=12, mod=11)
    out = net(x)
    if kind == "dqn":
        return env.decode_dqn(out.float().view(env.num_envs, 2, -1))
    idx = torch.multinomial(out.float(), 7, replacement=False)  # PPOAgent.get_action draws 7 distinct flat indices
    return env.decode_indices(idx.view(env.num_envs, 2, 7), div=12, mod=11)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs-per-gpu", type=int, default=32768)  # 262,144 over 8 GPUs
    ap.add_argument("--turns", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=150)
    ap.add_argument("--policy", default="dqn", choices=["dqn", "ppo", "rppo"])
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--graph", action="store_true", help="replay the turn (forward, decode, step) from a CUDA graph")
    ap.add_argument("--fused", action="store_true",
                    help="dqn only: the forward runs in libevgsim's fused tcgen05 kernel (evg_policy_mlp, bf16 operands) instead of torch")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    dtype = torch.float32 if args.dtype == "fp32" else torch.bfloat16
    E = args.envs_per_gpu
    first, _ = evgsim.shard_range(E * world, rank, world)
    env = evgsim.BatchedEvergladesEnv(E, device=local, seed=0, auto_reset=evgsim._capi.AUTORESET_TERMINAL, env_id_offset=first)
    net = build_policy(args.policy, dtype, env.device)
    env.reset()
    hidden = torch.zeros((1, 2 * E, 128), dtype=dtype, device=env.device) if args.policy == "rppo" else None

    fused = None
    if args.fused:
        assert args.policy == "dqn", "--fused is the DQN network's kernel"
        from evgsim import policy as evp
        fused = evp.FusedDQN(env, net.float())

    def turn():
        env.step(fused() if fused is not None else policy_actions(env, net, args.policy, dtype, hidden))
        if hidden is not None:  # a finished match starts over with a fresh hidden state
            hidden.mul_((1 - env.done).to(dtype).repeat_interleave(2).view(1, -1, 1))

    run = turn
    if args.graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                turn()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                turn()
        torch.cuda.current_stream().wait_stream(side)
        run = g.replay
    for _ in range(args.warmup):
        run()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.turns):
        run()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=env.device)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    stats = evgsim.gather_episode_stats(env.episode_stats(), device=env.device) if world > 1 else env.episode_stats()
    if rank == 0:
        sec = float(ms.item()) / 1e3
        print(json.dumps({"workload": "policy-in-the-loop self-play (BASELINE.json configs[3])", "policy": args.policy, "dtype": "bf16 operands, fp32 accumulate (evg_policy_mlp)" if args.fused else args.dtype, "fused_forward": bool(args.fused),
                          "cuda_graph": bool(args.graph),
                          "n_gpus": world, "envs_per_gpu": E, "turns": args.turns, "env_turns_per_s": E * world * args.turns / sec,
                          "ms_per_turn": sec * 1e3 / args.turns, "episodes": stats["episodes"], "wins": stats["wins"], "ties": stats["ties"]}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
