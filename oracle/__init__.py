"""CPU oracle for the k-mer spectrum hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product package (kmernator_b200/) never does.

* ``kmn_oracle.c``   -- C restatement of the reference algorithm (each function cites file:line).
* ``filter_oracle``  -- numpy/Python restatement of the FilterReads-level host logic (artifact
                        quality trim, labels, pair keep rule) used on small inputs.
* ``_ref/``          -- the reference's own ``src/lookup3.h`` compiled where it lies (hash pin).

Parity status: pinned against the reference's golden fixtures (tests/test_oracle_golden.py).
"""
from .binding import (  # noqa: F401
    OracleSpectrum,
    SCORING,
    SCORE_LABEL,
    concat_reads,
    build,
    compress_sequence,
    first_markup_n_or_x,
    histogram_bin,
    histogram_bucket_value,
    kmer_hash,
    kmer_hash_lookup8,
    lib,
    owner,
    passes_length,
    quality_table,
    read_kmers,
    ref_kmer_hash,
    reverse_complement,
    trim_values,
    estimate_raw_kmers,
)
