"""vafator_b200 -- B200-native implementation of VAFator's per-variant pileup-annotation hot path.

Drop-in surface (same names and argument meaning as TRON-Bioinformatics/vafator v3.0.0):
    vafator_b200.annotator.Annotator, vafator_b200.command_line.annotator (the `vafator` CLI),
    vafator_b200.pileups.collect_metrics_for_chrom, vafator_b200.rank_sum_test.get_rank_sum_tests,
    vafator_b200.power.PowerCalculator, vafator_b200.ploidies.PloidyManager.
All arithmetic of the path runs in libvfk.so (CUDA, sm_100a); there is no CPU fallback."""
VERSION = "3.0.0"  # the reference version whose INFO contract is reproduced (vafator/__init__.py:1)
