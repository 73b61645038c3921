"""Tumour ploidy: genome-wide value or local copy numbers from a BED file (mirror of
vafator/ploidies.py:11-53), with a vectorised lookup for whole variant lists."""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np

DEFAULT_PLOIDY = 2.0


class PloidyManager:

    def __init__(self, local_copy_numbers: str = None, genome_wide_ploidy: float = DEFAULT_PLOIDY):
        if local_copy_numbers is not None and not os.path.exists(local_copy_numbers):
            raise ValueError(
                "The provided tumor ploidy is neither a copy number value or a BED file with copy "
                "numbers"
            )
        self.report_value = local_copy_numbers if local_copy_numbers else genome_wide_ploidy
        self.ploidy = genome_wide_ploidy
        self.bed = None
        self._numeric_chroms = False
        if local_copy_numbers is not None:
            chroms: List[str] = []
            rows = []
            with open(local_copy_numbers) as fh:
                for line in fh:
                    if not line.strip():
                        continue
                    c, s, e, cn = line.rstrip("\n").split("\t")[:4]
                    chroms.append(c)
                    rows.append((float(s), float(e), float(cn)))
            arr = np.asarray(rows, np.float64).reshape(-1, 3)
            # The reference parses the BED with pandas.read_csv: a chromosome column that holds only
            # integers becomes an integer column and then never equals a (string) CHROM
            # (vafator/ploidies.py:26-31, :42-48) -- reproduced, since results must match.
            self._numeric_chroms = len(chroms) > 0 and all(_is_int(c) for c in chroms)
            self.bed = (np.asarray(chroms, object), arr[:, 0], arr[:, 1], arr[:, 2])

    def get_ploidy(self, variant) -> float:
        return float(self.get_ploidies([variant.CHROM], [variant.POS])[0]) if self.bed is not None else self.ploidy

    def get_ploidies(self, chroms: Sequence[str], positions: Sequence[int]) -> np.ndarray:
        """First BED row with start <= POS-1 < end on the same chromosome wins; default otherwise."""
        n = len(positions)
        out = np.full(n, float(self.ploidy))
        if self.bed is None or n == 0 or self._numeric_chroms:
            return out
        bchrom, bstart, bend, bcn = self.bed
        pos0 = np.asarray(positions, np.float64) - 1
        chroms = np.asarray(chroms, object)
        for c in set(bchrom.tolist()):
            rows = np.nonzero(bchrom == c)[0]
            sel = np.nonzero(chroms == c)[0]
            if len(sel) == 0:
                continue
            found = np.zeros(len(sel), bool)
            for r in rows:  # BED order: the first hit wins
                hit = ~found & (bstart[r] <= pos0[sel]) & (bend[r] > pos0[sel])
                out[sel[hit]] = bcn[r]
                found |= hit
        return out


def _is_int(text: str) -> bool:
    try:
        int(text)
        return True
    except ValueError:
        return False


default_ploidy_manager = PloidyManager()
