"""cProfile of the host-protocol loop (agent.step / environment.step / agent.update with numpy arrays)
at the bench shape: where do the ~400 us per vector step go?"""
import cProfile
import pstats
import sys
import time

sys.path.insert(0, '.')
import torch  # noqa: E402
import bench  # noqa: E402
from tonic_b200 import config  # noqa: E402
from tonic_b200.utils import logger  # noqa: E402

logger.store = lambda *a, **k: None
logger.store_aggregate = lambda *a, **k: None
config.noise = config.indices = 'device'
agent, env, batch = bench.build_ppo(bench.ENVS_PER_GPU, bench.ENVS_PER_GPU)
obs = env.start(host=True)
steps = 0


def iteration(obs, steps, n=bench.SEGMENT):
    for _ in range(n):
        actions = agent.step(obs, steps)
        obs, infos = env.step(actions)
        agent.update(**infos, steps=steps)
        steps += bench.ENVS_PER_GPU
    return obs, steps


for _ in range(4):
    obs, steps = iteration(obs, steps)
torch.cuda.synchronize()
t0 = time.time()
obs, steps = iteration(obs, steps, 127)      # no update inside: pure stepping
torch.cuda.synchronize()
print('us per vector step (no update):', (time.time() - t0) / 127 * 1e6)
obs, steps = iteration(obs, steps, 1)
prof = cProfile.Profile()
prof.enable()
obs, steps = iteration(obs, steps, 127)
prof.disable()
torch.cuda.synchronize()
st = pstats.Stats(prof)
st.sort_stats('tottime').print_stats(28)
st.sort_stats('cumulative').print_stats(22)
