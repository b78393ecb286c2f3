"""Off-policy ring replay resident in HBM (reference: tonic/replays/buffers.py).
Filled in with the off-policy agents."""


class Buffer:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('device ring replay: see round notes')
