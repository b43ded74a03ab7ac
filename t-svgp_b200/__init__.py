"""B200-native t-SVGP natural-gradient path (natgrad_step / elbo / predict_f over DenseSites)."""
from ._lib import (DeviceArray, pinned_empty, pinned_free, InvalidArgumentError, NonPositiveVarianceError, NotPositiveDefiniteError, TsvgpError, exported_names, load)
from .model import DenseSites, balance_weights, comm_unique_id, shard_rows, stream_minibatches, t_SVGP, t_SVGP_white

__all__ = ["DeviceArray", "pinned_empty", "pinned_free", "t_SVGP", "t_SVGP_white", "DenseSites", "comm_unique_id", "shard_rows", "balance_weights", "stream_minibatches", "load", "exported_names", "TsvgpError",
           "InvalidArgumentError", "NotPositiveDefiniteError", "NonPositiveVarianceError"]
