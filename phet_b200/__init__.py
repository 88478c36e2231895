"""phet_b200: the PHet ``fit_predict`` hot path of kleelab-bch/phet on NVIDIA B200 (sm_100a).

    from phet_b200 import PHet            # drop-in for ``from model.phet import PHet``
    scores = PHet(num_subsamples=1000).fit_predict(X, y)

Host side in Python (like the reference), device side in hand-written CUDA behind the C ABI of
``include/phet_b200.h`` (``phet_b200/libphet_b200.so``, built by ``python -m phet_b200.build``).
"""
from .phet import PHet  # noqa: F401
from .siblings import HIQR, DeltaIQRMean  # noqa: F401
from .ingest import StagedMatrix, load_mtx  # noqa: F401

__all__ = ["PHet", "HIQR", "DeltaIQRMean", "StagedMatrix", "load_mtx"]
