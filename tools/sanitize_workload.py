"""Workload for compute-sanitizer (tools/sanitize.sh): every kernel of libevgsim at small batches with real combat.

Scripted agents on the device (base_rushV1 vs SwarmAgent meet and fight from about turn 15 on; random_actions for
variety) drive 90 turns through each step kernel — warp per match, thread per match with 32- and with 128-thread
CTAs, the run-time-sized instantiation on another map — with auto-reset, the packed wire rows, masked resets and the
state export/import kernels.  Prints how many unit slots fought so the log shows the combat phase really ran."""
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import evgsim
from evgsim import _capi

turns = int(sys.argv[1]) if len(sys.argv) > 1 else 90


def drive(name, n, env_vars, cfg=None, agents=(_capi.AGENT_BASE_RUSH, _capi.AGENT_SWARM)):
    for k in ("EVG_STEP_KERNEL", "EVG_TPM_SMALL_MAX"):
        os.environ.pop(k, None)
    os.environ.update(env_vars)
    kw = {"config": cfg} if cfg is not None else {}
    env = evgsim.BatchedEvergladesEnv(n, seed=5, auto_reset=_capi.AUTORESET_NEXT, **kw)
    env.reset()
    for t in range(turns):
        if t % 3 == 2:
            a = env.random_actions()
        else:
            a = env.agent_actions(*agents)
        if t % 2:
            env.step(a, obs_format="wire")
        else:
            env.step(a)
        if t == 40:
            env.reset(mask=torch.arange(n, device=env.device) % 5 == 0)
    st = env.get_state()
    env.set_state(st)
    env.step_agents()
    env.shape_reward(_capi.SHAPE_SHORT_GAMES)
    torch.cuda.synchronize()
    s = env.episode_stats()
    print(json.dumps({"leg": name, "matches": n, "kernel_kind": env._lib.evg_step_kernel_kind(env._h), "turns": turns + 1,
                      "fought_unit_slots": s["fought_unit_slots"], "episodes": s["episodes"]}), flush=True)
    assert s["fought_unit_slots"] > 0, "no combat happened: the sanitizer run would prove nothing"
    env.close()


drive("warp-per-match", 96 + 5, {"EVG_STEP_KERNEL": "warp"})
drive("thread-per-match, 32-thread CTAs", 512 + 37, {"EVG_STEP_KERNEL": "tpm", "EVG_TPM_SMALL_MAX": str(1 << 30)})
drive("thread-per-match, 128-thread CTAs", 1024 + 37, {"EVG_STEP_KERNEL": "tpm", "EVG_TPM_SMALL_MAX": "0"})
from test_gpu_generic import write_configs
import pathlib
d = write_configs(pathlib.Path(tempfile.mkdtemp()), 120)
cfg7 = evgsim.load_config(d, "Map7.json", "Units4.json", "Setup.json", auto_reset=_capi.AUTORESET_NEXT)
drive("thread-per-match, run-time-sized map", 256 + 11, {"EVG_STEP_KERNEL": "tpm"}, cfg=cfg7, agents=(_capi.AGENT_RANDOM, _capi.AGENT_RANDOM))
print("sanitize workload done")
