"""Power model (mirror of vafator/power.py:14-131).  Same class, same method names and argument
meaning; the binomial arithmetic runs in the FP64 epilogue kernel (vfk_epilogue), the expected
VAF (a closed formula over purity / ploidy) on the host as in the reference."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .ploidies import default_ploidy_manager

DEFAULT_PURITY = 1.0
DEFAULT_NORMAL_PLOIDY = 2
DEFAULT_FPR = 5 * (10 ** -7)
DEFAULT_ERROR_RATE = 10 ** -3


class PowerCalculator:

    def __init__(self, tumor_ploidies: dict, purities: dict, normal_ploidy: int = DEFAULT_NORMAL_PLOIDY,
                 fpr: float = DEFAULT_FPR, error_rate: float = DEFAULT_ERROR_RATE, device_index: int = 0):
        self.normal_ploidy = normal_ploidy
        self.purities = purities
        self.tumor_ploidies = tumor_ploidies
        self.fpr = fpr
        self.error_rate = error_rate
        self._device_index = device_index
        self._engine = None

    # ---- expected VAF (vafator/power.py:52-82) ----
    def calculate_expected_vaf(self, sample: str, variant) -> float:
        purity = self.purities.get(sample, DEFAULT_PURITY)
        tumor_ploidy = max(1, self.tumor_ploidies.get(sample, default_ploidy_manager).get_ploidy(variant=variant)
                           if variant is not None else self.tumor_ploidies.get(sample, default_ploidy_manager).ploidy)
        corrected = purity * tumor_ploidy + ((1 - purity) * self.normal_ploidy)
        return purity / corrected

    def expected_vafs(self, sample: str, chroms: Sequence[str], positions: Sequence[int]) -> np.ndarray:
        """Vectorised calculate_expected_vaf over a list of variants."""
        purity = self.purities.get(sample, DEFAULT_PURITY)
        ploidy = np.maximum(1, self.tumor_ploidies.get(sample, default_ploidy_manager).get_ploidies(chroms, positions))
        return purity / (purity * ploidy + ((1 - purity) * self.normal_ploidy))

    # ---- device-backed statistics ----
    def _stats(self, dp: int, ac: int, eaf: float):
        from . import device as _dev
        from .engine import Engine
        if self._engine is None:
            self._engine = Engine(self._device_index, fpr=self.fpr, error_rate=self.error_rate)
        c = np.zeros(1, _dev.COUNTS_DTYPE)
        c["dp"], c["ac_alt"] = dp, ac
        return self._engine.epilogue(c, np.asarray([eaf], np.float64))[0]

    def calculate_power(self, dp: int, ac: int, sample: str, variant) -> float:
        """P[X <= ac], X ~ Binomial(dp, expected VAF), rounded to 5 decimals (vafator/power.py:39-50)."""
        st = self._stats(dp, ac, self.calculate_expected_vaf(sample, variant))
        return float(np.round(np.float64(st["pu"]), 5))

    def calculate_absolute_power(self, sample, variant, dp: int):
        """(Carter-2012 power rounded to 5 decimals, k) -- vafator/power.py:116-131."""
        st = self._stats(dp, 0, self.calculate_expected_vaf(sample, variant))
        return float(np.round(np.float64(st["pw"]), 5)), int(st["k"])
