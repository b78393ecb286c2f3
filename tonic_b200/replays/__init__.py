from .buffers import Buffer
from .segments import Segment
from .utils import flatten_batch, lambda_returns

__all__ = [flatten_batch, lambda_returns, Buffer, Segment]
