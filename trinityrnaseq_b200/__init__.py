"""trinityrnaseq_b200 -- B200-native k-mer hot path of Trinity (count -> coverage stats -> read assignment).

Python here is only the host-side mirror of the reference's interfaces for tests and bench; the product is
libtrinity_gpu.so (CUDA, sm_100a) plus the drop-in executables in trinityrnaseq_b200/bin/.
"""
from .api import (Context, KmerCounter, BundleKmerTable, WeldmerTable, records_from_sequences,  # noqa: F401
                  format_stats_line, packed_to_kmer, kmer_to_packed)
from ._lib import TrinityGpuError  # noqa: F401
from . import sharded  # noqa: F401
