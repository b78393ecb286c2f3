"""Command line with the reference's flags (tonic/train.py:140-159):

    python -m tonic_b200.train --header "import tonic_b200.torch" \
        --agent "tonic_b200.torch.agents.PPO()" \
        --environment "tonic_b200.environments.SynthControl('HalfCheetah')" \
        --parallel 1 --sequential 4096 --seed 0

`--parallel P --sequential S` create P*S environments in total (sharded over the
ranks when launched with torchrun); header / agent / environment / trainer are
Python snippets evaluated like in the reference (train.py:79-137).  The name
`tonic` is bound to this package, so reference command lines
(`tonic.torch.agents.PPO()`, `tonic.Trainer(...)`) work unchanged apart from the
environment string.  Checkpoint resumption (`--path`, `--checkpoint`) follows
train.py:22-75,101-103.
"""

import argparse
import os

import yaml

import tonic_b200
import tonic_b200.torch  # noqa: F401

tonic = tonic_b200       # reference-style snippets (`tonic.torch.agents.PPO()`) see `tonic`


def train(header, agent, environment, test_environment, trainer, before_training,
          after_training, parallel, sequential, seed, name, environment_name, checkpoint, path):
    args = dict(locals())
    checkpoint_path = None
    if path:
        tonic_b200.logger.log(f'Loading experiment from {path}')
        if not (checkpoint == 'none' or agent is not None):
            folder = os.path.join(path, 'checkpoints')
            ids = [int(f.split('.')[0][5:]) for f in os.listdir(folder) if f.startswith('step_')] \
                if os.path.isdir(folder) else []
            if not ids:
                tonic_b200.logger.error(f'No checkpoint found in {folder}')
            elif checkpoint == 'last':
                checkpoint_path = os.path.join(folder, f'step_{max(ids)}')
            elif int(checkpoint) in ids:
                checkpoint_path = os.path.join(folder, f'step_{int(checkpoint)}')
            else:
                tonic_b200.logger.error(f'Checkpoint {checkpoint} not found in {folder}')
        with open(os.path.join(path, 'config.yaml')) as config_file:
            config = argparse.Namespace(**yaml.load(config_file, Loader=yaml.FullLoader))
        header = header or config.header
        agent = agent or config.agent
        environment = environment or config.test_environment or config.environment
        trainer = trainer or config.trainer

    if header:
        exec(header)

    _environment = environment
    environment = tonic_b200.environments.distribute(
        lambda: eval(_environment), parallel, sequential)
    environment.initialize(seed=seed)

    _test_environment = test_environment if test_environment else _environment
    # a single test environment built the way the reference builds it (train.py:88-91):
    # distribute() picks the device grid or the host grid from what the builder returns
    test_environment = tonic_b200.environments.distribute(
        lambda: eval(_test_environment), single=True)
    test_environment.initialize(seed=seed + 10000)

    if not agent:
        raise ValueError('No agent specified.')
    agent = eval(agent)
    agent.initialize(observation_space=environment.observation_space,
                     action_space=environment.action_space, seed=seed)
    if checkpoint_path:
        agent.load(checkpoint_path)

    environment_name = environment_name or getattr(test_environment, 'name',
                                                   test_environment.__class__.__name__)
    if not name:
        name = getattr(agent, 'name', agent.__class__.__name__)
        if parallel != 1 or sequential != 1:
            name += f'-{parallel}x{sequential}'
    tonic_b200.logger.initialize(os.path.join(environment_name, name, str(seed)),
                                 script_path=__file__, config=args)

    trainer = eval(trainer or 'tonic_b200.Trainer()')
    trainer.initialize(agent=agent, environment=environment, test_environment=test_environment)
    if before_training:
        exec(before_training)
    trainer.run()
    if after_training:
        exec(after_training)
    return trainer


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--header')
    parser.add_argument('--agent')
    parser.add_argument('--environment', '--env')
    parser.add_argument('--test_environment', '--test_env')
    parser.add_argument('--trainer')
    parser.add_argument('--before_training')
    parser.add_argument('--after_training')
    parser.add_argument('--parallel', type=int, default=1)
    parser.add_argument('--sequential', type=int, default=1)
    parser.add_argument('--seed', type=int, default=0)
    parser.add_argument('--name')
    parser.add_argument('--environment_name')
    parser.add_argument('--checkpoint', default='last')
    parser.add_argument('--path')
    train(**vars(parser.parse_args()))


if __name__ == '__main__':
    main()
