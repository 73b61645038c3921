"""Host logic on CPU: the composition layer (vafator_b200/engine.py: replicate merge, Counter / dict
lookup semantics, rounding, INFO ordering) driven by ORACLE counts instead of the device, against
the INFO strings the unmodified reference produced (tests/golden).  The device stages are replaced
by test doubles here only -- the product's Engine always calls libvfk.so."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import scipy.special

import helpers as H
from oracle import c_oracle, vafator_ref as vr
from vafator_b200 import device, engine

SCENARIOS = H.golden_scenarios()


class OracleEngine(engine.Engine):
    def __init__(self, min_mapq, min_baseq, include_ambiguous, fpr, error_rate, normal_ploidy=2):
        self.params = device.make_params(min_mapq=min_mapq, min_baseq=min_baseq, include_ambiguous=include_ambiguous)
        self.fpr, self.error_rate = fpr, error_rate
        self._prm = (min_mapq, min_baseq, include_ambiguous)
        self._records = None

    def pileup(self, batch, units, stop):
        recs = self._records
        ou = [(batch.contigs.index(recs[i][0]), recs[i][1], recs[i][2], recs[i][3][j], recs[i][3][0], int(s))
              for i, j, s in zip(units.rec, units.alt_index, stop)]
        o = c_oracle.pileup(batch.as_dict(), ou, c_oracle.params(min_mapq=self._prm[0], min_baseq=self._prm[1]))
        c = np.zeros(len(ou), device.COUNTS_DTYPE)
        snv = (units.meta & 3) == 0
        c["dp"] = o["dp"] - (0 if self._prm[2] else np.where(snv, o["n_amb"], 0))
        for f in ("n_amb", "ac_ref", "ac_alt", "med2", "s2", "n_col"):
            c[f] = o[f]
        c["status"] = units.meta & 3
        return c

    def values(self, batch, units, stop):
        """Test double of vfk_pileup_values: the per-read lists from the Python restatement."""
        from oracle import htslib_pileup as hp
        recs = self._records
        f = hp.AlignmentFile(hp.recs_from_soa(batch.as_dict()), batch.contigs)
        out = []
        for u in range(units.n_units):
            chrom, pos1 = batch.contigs[units.tid[u]], int(units.pos[u]) + 1
            kind = int(units.meta[u]) & 3
            assert kind == 0
            ref_code, alt_code = (int(units.meta[u]) >> 8) & 0xFF, (int(units.meta[u]) >> 16) & 0xFF
            words = []
            for col in f.pileup(chrom, pos1 - 1, pos1, min_base_quality=self._prm[1], min_mapping_quality=self._prm[0],
                                max_depth=10 ** 6):
                for pr in col.pileups:
                    if pr.is_del or pr.is_refskip:
                        continue
                    q = pr.query_position
                    code = hp.SEQ_CHARS.index(pr.alignment.query_sequence[q])
                    w = pr.alignment.mapping_quality | (pr.alignment.query_qualities[q] << 8) | (q << 16)
                    w |= (1 << 28) if code == ref_code else 0
                    w |= (1 << 29) if code == alt_code else 0
                    if w >> 28:
                        words.append(w)
            out.append(np.asarray(words, np.uint32))
        return out

    def ranksum_s2(self, alt_lists, ref_lists):
        out = np.zeros((len(alt_lists), 3), np.uint64)
        for p, (a, r) in enumerate(zip(alt_lists, ref_lists)):
            for m, (sh, mk) in enumerate(((0, 0xFF), (8, 0xFF), (16, 0xFFF))):
                out[p, m] = H.midrank_s2([(int(x) >> sh) & mk for x in a], [(int(x) >> sh) & mk for x in r])
        return out

    def epilogue(self, counts, eaf):
        st = np.zeros(len(counts), device.STATS_DTYPE)
        power = vr.Power(fpr=self.fpr, error_rate=self.error_rate)
        for u, c in enumerate(counts):
            n1, n2 = int(c["ac_alt"]), int(c["ac_ref"])
            for m in range(3):
                if n1 and n2 and (m != 1 or (int(c["status"]) & 3) == 0):
                    z = (float(c["s2"][m]) / 2.0 - n1 * (n1 + n2 + 1) / 2.0) / np.sqrt(n1 * n2 * (n1 + n2 + 1) / 12.0)
                    st["z"][u][m], st["p"][u][m] = z, 2 * scipy.special.ndtr(-abs(z))
                else:
                    st["z"][u][m] = st["p"][u][m] = np.nan
            st["pu"][u] = power.pu_raw(int(c["dp"]), n1, eaf[u])
            st["pw"][u], st["k"][u] = power.pw_raw(int(c["dp"]), eaf[u])
        return st


@pytest.mark.parametrize("path", SCENARIOS, ids=[os.path.basename(p) for p in SCENARIOS])
def test_info_strings_match_reference(path):
    sc = H.load_scenario(path)
    records = [(v["chrom"], v["pos"], v["ref"], v["alts"]) for v in sc["variants"]]
    batches = H.scenario_batches(sc)
    for run in sc["runs"]:
        prm = run["params"]
        eng = OracleEngine(prm["min_mq"], prm["min_bq"], prm["include_ambiguous"], prm.get("fpr", 5e-7),
                           prm.get("error_rate", 1e-3))
        eng._records = records
        pl = {s: ([tuple(r) for r in p] if isinstance(p, list) else p) for s, p in sc["ploidies"].items()}
        power = vr.Power(ploidy=pl, purity=dict(sc["purities"]), normal_ploidy=prm.get("normal_ploidy", 2))
        eaf = {s: np.asarray([power.eaf(s, r[0], r[1]) for r in records]) for s in sc["samples"]}
        bams = OrderedDict((s, [batches[b] for b in idx]) for s, idx in sc["samples"].items())
        got = eng.annotate(records, bams, eaf)
        for i, (g, want) in enumerate(zip(got, run["info"])):
            w = OrderedDict((k, v) for k, v in want)
            assert list(g.keys()) == list(w.keys()), (sc["name"], prm, records[i], list(g.keys()), list(w.keys()))
            for k in w:
                assert str(g[k]) == str(w[k]), (sc["name"], prm, records[i], k, g[k], w[k])
