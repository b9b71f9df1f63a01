"""gretel_b200 - B200-native Hansel matrix ingestion and Gretel haplotype recovery.

Only the hot path of SamStudio8/gretel lives here:
  gretel_b200.hansel.Hansel                      drop-in for hansel.Hansel
  gretel_b200.util.load_from_bam / process_vcf   drop-in for gretel.util
  gretel_b200.gretel.generate_path / reweight_hansel_from_path   drop-in for gretel.gretel
The compute lives in libhanselx.so (hand-written sm_100a CUDA behind the C ABI of
include/hanselx.h).  Importing this package does not load the library; the first
Hansel does, and raises if it is missing.
"""
__version__ = "0.1.0"
