from .mean_stds import MeanStd

__all__ = [MeanStd]
