"""Helpers for the -m gpu parity tests: run the CUDA path through the C ABI and the oracle on the same bytes."""
import numpy as np

import oracle
from oracle import filter_oracle as F


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def oracle_table(bases, quals, off, k, disc=None, start=33, min_quality=3, min_kmer_quality=0.10, track_ext=False, threads=1):
    s = oracle.OracleSpectrum(k, min_quality, min_kmer_quality, track_ext=track_ext, threads=threads,
                              est_distinct=max(1024, len(bases) // 8), start=start)
    s.add_reads(bases if isinstance(bases, bytes) else bases.tobytes(), quals, off, disc)
    return s


def assert_tables_equal(g, o, check_dir=True, check_wsum=False, check_ext=False):
    assert g["keys"].shape == o["keys"].shape, (g["keys"].shape, o["keys"].shape)
    assert (g["keys"] == o["keys"]).all()
    assert (g["count"] == o["count"]).all()
    if check_dir:
        # directionBias: the reference loses the strand of the first observation on singleton promotion
        # (KmerTrackingData.h:654-656), so it is order-dependent by +-1 in the reference itself.
        d = g["dir"].astype(np.int64) - o["dir"].astype(np.int64)
        assert ((d == 0) | (d == 1)).all()
    if check_wsum:
        assert np.allclose(g["wsum"], o["wsum"], rtol=1e-3, atol=5e-3)
    if check_ext:
        assert (g["ext"] == o["ext"]).all()
