"""One PPO iteration with a single 16384-row minibatch per network (4096 envs x T=4, E=1),
graphs off, bracketed by cudaProfilerStart/Stop so that
`ncu --profile-from-start off --set full` captures exactly one launch of every kernel of
the update chain at the bench's minibatch size."""
import sys

import torch

sys.path.insert(0, '.')
from oracle import scenarios  # noqa: E402  (config dictionary only)
from tests import product  # noqa: E402
from tonic_b200 import config  # noqa: E402

config.noise = config.indices = 'device'
config.graphs = False
cfg = dict(scenarios.SCENARIOS['ppo_wide'], obs=17, act=6, workers=4096, hidden=(256, 256),
           max_episode_steps=1000, segment=dict(size=4, batch_iterations=1, batch_size=16384))
agent, env = product.build(cfg)
env.start()
for _ in range(3):
    agent.rollout(env, 4)
torch.cuda.synchronize()
torch.cuda.profiler.start()
agent.rollout(env, 4)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done')
