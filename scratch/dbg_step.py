"""Debug aid: the fused-step equality test's scenario, one variant per process, with a watchdog."""
import faulthandler
import sys

faulthandler.dump_traceback_later(45, exit=True)
sys.path.insert(0, '.')
import torch  # noqa: E402
from oracle import scenarios  # noqa: E402
from tests import product  # noqa: E402
from tonic_b200 import config  # noqa: E402

fused = sys.argv[1] == '1'
kind = sys.argv[2] if len(sys.argv) > 2 else 'PPO'
base = scenarios.SCENARIOS['ppo_wide' if kind == 'PPO' else 'a2c_small']
seg = dict(base['segment'], size=24)
cfg = dict(base, workers=100, max_episode_steps=9, segment=seg)
config.noise, config.indices, config.graphs = 'device', 'device', False
config.fused_rollout, config.fused_step = False, fused
agent, env = product.build(cfg)
env.start()
torch.cuda.synchronize()
print('built', flush=True)
for k in range(10):
    assert agent.rollout(env, 1) == 1
    torch.cuda.synchronize()
    print('step', k, flush=True)
print('rollout 2', flush=True)
n = agent.rollout(env, seg['size'])
torch.cuda.synchronize()
print('done', n, int(agent._noise_counter.item()), flush=True)
