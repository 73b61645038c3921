#!/usr/bin/env python
"""ORACLE tooling (test infrastructure) -- generates tests/golden/*.json by running the
UNMODIFIED reference code from /root/reference on top of oracle/htslib_pileup.py.

What is genuinely the reference here: vafator/pileups.py (collect_metrics_for_chrom and the
three per-column functions), vafator/annotator.py (_add_stats / _annotate_sample /
_annotate_replicate / _calculate_af), vafator/rank_sum_test.py, vafator/power.py,
vafator/ploidies.py -- imported as they lie, with a stub for the two absent third-party
modules.  What is restated: the pysam/htslib objects those functions
read (oracle/htslib_pileup.py, parity unpinned -- see its header).

Run only in the build container (needs /root/reference):
    python oracle/gen_golden.py
"""
import json
import os
import random
import sys
import types
from collections import OrderedDict

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

REF_ROOT = os.environ.get("VAFATOR_REFERENCE", "/root/reference")


def import_reference():
    cy = types.ModuleType("cyvcf2")
    cy.Variant = cy.VCF = cy.Writer = object
    sys.modules["cyvcf2"] = cy
    sys.modules["pysam"] = types.ModuleType("pysam")
    pl = types.ModuleType("pysam.libcalignmentfile")
    pl.IteratorColumnRegion = pl.AlignmentFile = object
    sys.modules["pysam.libcalignmentfile"] = pl
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import vafator.pileups as rp
    import vafator.annotator as ra
    import vafator.power as rw
    import vafator.ploidies as rl
    import vafator.rank_sum_test as rr
    return rp, ra, rw, rl, rr


from oracle import htslib_pileup as hp  # noqa: E402

BASES = "ACGT"
QUALS = (2, 12, 23, 30, 37, 41)
MAPQS = (0, 1, 10, 20, 29, 40, 60, 60, 60, 60, 60, 60, 60)


class FakeVariant:
    """Quacks like cyvcf2.Variant for the attributes annotator.py touches."""

    def __init__(self, chrom, pos, ref, alts):
        self.CHROM, self.POS, self.REF, self.ALT = chrom, pos, ref, list(alts)
        self.INFO = OrderedDict()


# ---------------------------------------------------------------------------
# scenario construction
# ---------------------------------------------------------------------------
def rec_dict(qname, flag, tid, pos, mapq, cigar, seq, qual, mtid=-1, mpos=-1, isize=0):
    assert len(seq) == len(qual)
    return dict(qname=qname, flag=flag, tid=tid, pos=pos, mapq=mapq, cigar=cigar, seq=seq,
                qual="".join(chr(q + 33) for q in qual), mtid=mtid, mpos=mpos, isize=isize)


def to_rec(d, index=-1):
    return hp.Rec(d["qname"], d["flag"], d["tid"], d["pos"], d["mapq"], d["cigar"], d["seq"],
                  bytes(ord(c) - 33 for c in d["qual"]), d["mtid"], d["mpos"], d["isize"], index=index)


def handmade():
    """By-inspection cases; one contig 'chrT' (tid 0).  Reference-ish context:
       pos(0-based) 100.. : ACGTTTTTGCA..."""
    R = []
    q = lambda n, v=37: [v] * n
    # --- overlapping mates, same base, Q20+Q20 -> one Q40, one Q0 (column 109, 1-based 110)
    R.append(rec_dict("pairA", 99, 0, 100, 60, "20M", "ACGTTTTTGCACGTACGTAC", q(20, 20), 0, 105, 25))
    R.append(rec_dict("pairA", 147, 0, 105, 60, "20M", "TTTGCACGTACGTACGGGGG", q(20, 20), 0, 100, -25))
    # --- overlapping mates that disagree at column 112 (0-based): Q37 vs Q37 -> 29 / 0
    R.append(rec_dict("pairB", 99, 0, 101, 60, "20M", "CGTTTTTGCACATACGTACG", q(20), 0, 106, 25))
    R.append(rec_dict("pairB", 147, 0, 106, 60, "20M", "TTGCACGTACGTACGGGGGA", q(20), 0, 101, -25))
    # --- disagreeing mates with different qualities (higher one keeps 0.8*q)
    R.append(rec_dict("pairC", 83, 0, 102, 60, "20M", "GTTTTTGCACTTACGTACGG", q(20, 41), 0, 107, 25))
    R.append(rec_dict("pairC", 163, 0, 107, 60, "20M", "TGCACGTACGTACGGGGGAA", q(20, 30), 0, 102, -25))
    # --- soft clips both ends, insertion in the T homopolymer (103..107), deletion, ref skip
    R.append(rec_dict("clip1", 0, 0, 103, 60, "3S10M4S", "GGGTTTTTGCACGAAAA", q(17)))
    R.append(rec_dict("ins1", 0, 0, 100, 29, "6M2I12M", "ACGTTTTTTTGCACGTACGT", q(20)))
    R.append(rec_dict("ins2", 16, 0, 100, 60, "2S6M2I10M", "NNACGTTTTTTTGCACGTAC", q(20)))
    R.append(rec_dict("ins3", 0, 0, 100, 60, "6M3I11M", "ACGTTTTTTTTGCACGTACG", q(20)))
    R.append(rec_dict("ins4", 0, 0, 100, 60, "6M2I2I10M", "ACGTTTTTTTTTGCACGTAC", q(20)))  # consecutive I
    R.append(rec_dict("ins5", 0, 0, 100, 60, "6M2I1P1I11M", "ACGTTTAATTTGCACGTACG", q(20)))  # pad
    R.append(rec_dict("del1", 0, 0, 100, 60, "8M2D12M", "ACGTTTTTACGTACGTACGG", q(20)))
    R.append(rec_dict("del2", 16, 0, 100, 40, "8M1D1D11M", "ACGTTTTTACGTACGTACG", q(19)))  # 1D1D merges to 2D
    R.append(rec_dict("del3", 0, 0, 100, 60, "8M3D12M", "ACGTTTTTCGTACGTACGGG", q(20)))
    R.append(rec_dict("skip1", 0, 0, 100, 60, "5M10N10M", "ACGTTACGTACGGGG", q(15)))
    R.append(rec_dict("deltail", 0, 0, 100, 60, "8M2D1I4M", "ACGTTTTTAACGT", q(13)))  # D followed by I
    # --- flags: dup / secondary / qcfail filtered; supplementary kept; orphan dropped; unpaired kept
    R.append(rec_dict("dup1", 1024, 0, 100, 60, "20M", "ACGTTTTTGCACGTACGTAC", q(20)))
    R.append(rec_dict("sec1", 256, 0, 100, 60, "20M", "ACGTTTTTGCACGTACGTAC", q(20)))
    R.append(rec_dict("qcf1", 512, 0, 100, 60, "20M", "ACGTTTTTGCACGTACGTAC", q(20)))
    R.append(rec_dict("sup1", 2048, 0, 100, 60, "20M", "ACGTTTTTGCACGTACGTAC", q(20)))
    R.append(rec_dict("orph", 73, 0, 100, 60, "20M", "ACGTTTTTGCACGTACGTAC", q(20), 0, 100, 0))
    R.append(rec_dict("lowmq", 0, 0, 100, 0, "20M", "ACGTTTTTGCACGTACGTAC", q(20)))
    # --- IUPAC / N / '=' bases and low qualities at column 110 (0-based)
    R.append(rec_dict("iupac", 0, 0, 104, 60, "16M", "TTTTGCNCGTRCGTAC", q(16)))
    R.append(rec_dict("eqbase", 0, 0, 104, 60, "16M", "TTTTGC=CGTACGTAC", q(16)))
    R.append(rec_dict("lowbq", 0, 0, 104, 60, "16M", "TTTTGCACGTACGTAC", [37] * 6 + [5] + [37] * 9))
    # --- supplementary alignment sharing a name with a pair (three-way hash protocol)
    R.append(rec_dict("tri", 99, 0, 108, 60, "12M", "CACGTACGTACG", q(12, 30), 0, 112, 16))
    R.append(rec_dict("tri", 2048 | 99, 0, 110, 60, "4S8M", "GGGGCGTACGTA", q(12, 30), 0, 112, 16))
    R.append(rec_dict("tri", 147, 0, 112, 60, "12M", "TACGTACGGGGG", q(12, 30), 0, 108, -16))
    # --- second contig, read ending in a deletion-adjacent column, zero-length alignment
    R.append(rec_dict("c2a", 0, 1, 10, 60, "10M", "ACGTACGTAC", q(10)))
    R.append(rec_dict("c2b", 16, 1, 12, 60, "4M2D4M", "GTACACGT", q(8)))
    R.append(rec_dict("zero", 0, 1, 14, 60, "6S", "AAAAAA", q(6)))
    R.sort(key=lambda d: (d["tid"], d["pos"]))
    V = [
        ("chrT", 110, "C", ["A"]), ("chrT", 110, "C", ["T", "A"]), ("chrT", 113, "G", ["A"]),
        ("chrT", 112, "A", ["T"]), ("chrT", 111, "a", ["c"]), ("chrT", 111, "A", ["N", "R", "="]),
        ("chrT", 106, "T", ["TTT"]), ("chrT", 106, "T", ["TTTT"]), ("chrT", 106, "T", ["TAA"]),
        ("chrT", 106, "T", ["TTT", "TTTT"]), ("chrT", 106, "T", ["Ttt"]), ("chrT", 106, "t", ["TTT"]),
        ("chrT", 108, "TGC", ["T"]), ("chrT", 108, "TAA", ["T"]), ("chrT", 108, "TGCA", ["T"]),
        ("chrT", 108, "TGC", ["t"]), ("chrT", 109, "G", ["C"]), ("chrT", 105, "T", ["G"]),
        ("chrT", 105, "TTTT", ["GGGG"]), ("chrT", 105, "TT", ["GGG"]), ("chrT", 120, "C", ["A", "GT"]),
        ("chrT", 300, "A", ["C"]), ("chrT", 119, "A", ["C"]),
        ("chrU", 13, "G", ["T"]), ("chrU", 15, "CAC", ["C"]), ("chrU", 16, "A", ["C"]), ("chrU", 17, "C", ["T"]),
    ]
    V.sort(key=lambda v: (("chrT", "chrU").index(v[0]), v[1]))
    return dict(name="handmade", contigs=["chrT", "chrU"], bams=[R],
                variants=[dict(chrom=c, pos=p, ref=r, alts=a) for c, p, r, a in V],
                samples=OrderedDict([("tumor", [0])]), purities={}, ploidies={})


def simulate_bam(rng, ref, events, n_pairs, read_len, tag):
    """Paired reads over `ref` carrying the shared `events` ({pos0: (kind, payload, vaf)})."""
    G = len(ref)
    out = []

    def build(start, strand_flag, name, mate_start, isize, extra_flag=0):
        # walk the sample genome from `start`, consuming read_len query bases
        ops, seq = [], []
        r = start
        lead = rng.choice([0, 0, 0, 0, 2, 5]) if rng.random() < 0.15 else 0
        trail = rng.choice([0, 0, 0, 3, 6]) if rng.random() < 0.15 else 0
        body = read_len - lead - trail
        if lead:
            ops.append(("S", lead))
            seq += [rng.choice(BASES) for _ in range(lead)]

        def add(op, n):
            if ops and ops[-1][0] == op:
                ops[-1] = (op, ops[-1][1] + n)
            else:
                ops.append((op, n))

        used = 0
        while used < body and r < G:
            ev = events.get(r)
            base = ref[r]
            carried = ev is not None and rng.random() < ev[2]
            if carried and ev[0] == "snv":
                base = ev[1]
            if rng.random() < 0.01:
                base = rng.choice(BASES)
            if rng.random() < 0.004:
                base = rng.choice("NNNRYK")
            add("M", 1)
            seq.append(base)
            used += 1
            r += 1
            if used >= body or r >= G:
                break
            if carried and ev[0] == "ins" and used + len(ev[1]) < body:
                add("I", len(ev[1]))
                seq += list(ev[1])
                used += len(ev[1])
            elif carried and ev[0] == "del" and r + ev[1] < G:
                add("D", ev[1])
                r += ev[1]
            elif rng.random() < 0.002 and used + 3 < body:  # private noise indels / skips
                kind = rng.choice("IDN")
                n = rng.randint(1, 3)
                if kind == "I":
                    add("I", n)
                    seq += [rng.choice(BASES) for _ in range(n)]
                    used += n
                elif r + n + 1 < G:
                    add(kind, n if kind == "D" else n + 7)
                    r += n if kind == "D" else n + 7
        if ops and ops[-1][0] in "DN":  # never end the aligned part on D/N
            ops.pop()
        if trail:
            ops.append(("S", trail))
            seq += [rng.choice(BASES) for _ in range(trail)]
        seq = "".join(seq)
        cigar = "".join("%d%s" % (n, o) for o, n in ops)
        qual = [rng.choice(QUALS) if rng.random() < 0.7 else 37 for _ in seq]
        return rec_dict(name, strand_flag | extra_flag, 0, start, rng.choice(MAPQS), cigar, seq, qual, 0, mate_start, isize)

    for i in range(n_pairs):
        frag = max(read_len // 2 + 5, int(rng.gauss(1.6 * read_len, 0.45 * read_len)))
        s1 = rng.randint(0, G - frag - 2 * read_len)
        s2 = max(s1, s1 + frag - read_len)
        name = "%s_%05d" % (tag, i)
        u = rng.random()
        if u < 0.80:
            f1, f2 = (99, 147) if rng.random() < 0.5 else (163, 83)
            x = 0
            v = rng.random()
            if v < 0.03:
                x = 1024
            elif v < 0.04:
                x = 256
            elif v < 0.05:
                x = 512
            elif v < 0.07:
                x = 2048
            out.append(build(s1, f1, name, s2, frag, x if rng.random() < 0.5 else 0))
            out.append(build(s2, f2, name, s1, -frag, x if rng.random() < 0.5 else 0))
        elif u < 0.86:  # orphans: paired, not proper
            out.append(build(s1, 65, name, s2, frag))
            out.append(build(s2, 129, name, s1, -frag))
        elif u < 0.90:  # mate unmapped but flagged proper (malformed but seen in the wild)
            out.append(build(s1, 99 | 8, name, s1, 0))
        else:  # single-end
            out.append(build(s1, rng.choice([0, 16]), name, -1, 0))
            out[-1]["mtid"] = -1
    out.sort(key=lambda d: (d["tid"], d["pos"]))
    return out


def random_scenario(seed, n_bams=1, n_pairs=260, read_len=60, G=1500, replicates=False):
    rng = random.Random(seed)
    ref = "".join(rng.choice(BASES) for _ in range(G))
    # make some homopolymer / tandem context so that the reference's shifted insertion
    # comparison can succeed
    ref = list(ref)
    for _ in range(12):
        p = rng.randint(50, G - 60)
        b = rng.choice(BASES)
        for j in range(rng.randint(4, 9)):
            ref[p + j] = b
    ref = "".join(ref)
    events, variants = {}, []
    for _ in range(70):
        p = rng.randint(40, G - 80)
        if any(abs(p - e) < 6 for e in events):
            continue
        u = rng.random()
        if u < 0.6:
            alt = rng.choice([b for b in BASES if b != ref[p]])
            events[p] = ("snv", alt, rng.choice([0.1, 0.3, 0.5, 1.0]))
            alts = [alt]
            if rng.random() < 0.2:
                alts.append(rng.choice([b for b in BASES if b not in (ref[p], alt)]))
            variants.append((p + 1, ref[p], alts))
        elif u < 0.8:
            n = rng.randint(1, 4)
            ins = ref[p + 1:p + 1 + n] if rng.random() < 0.7 else "".join(rng.choice(BASES) for _ in range(n))
            events[p] = ("ins", ins, rng.choice([0.3, 0.5, 1.0]))
            variants.append((p + 1, ref[p], [ref[p] + ins]))
        else:
            n = rng.randint(1, 5)
            events[p] = ("del", n, rng.choice([0.3, 0.5, 1.0]))
            variants.append((p + 1, ref[p:p + 1 + n], [ref[p]]))
    # decoys: positions with no event, MNV, wrong-length indels, uncovered end
    for _ in range(15):
        p = rng.randint(40, G - 80)
        variants.append((p + 1, ref[p], [rng.choice([b for b in BASES if b != ref[p]])]))
    variants.append((200, ref[199:201], ["GG"]))
    variants.append((G - 1, ref[G - 2], ["A" if ref[G - 2] != "A" else "C"]))
    variants.sort(key=lambda v: v[0])
    bams = [simulate_bam(rng, ref, events, n_pairs, read_len, "b%d" % b) for b in range(n_bams)]
    if replicates:
        samples = OrderedDict([("tumor", list(range(n_bams - 1))), ("normal", [n_bams - 1])])
        purities = {"tumor": 0.7}
        ploidies = {"tumor": [["chr1", 0, 500, 3.2], ["chr1", 500, 900, 0.6], ["chr1", 800, 1200, 5.0]],
                    "normal": 2.0}
    else:
        samples = OrderedDict(("s%d" % b, [b]) for b in range(n_bams))
        purities, ploidies = {}, {}
    return dict(name="random_%d" % seed, contigs=["chr1"], bams=bams,
                variants=[dict(chrom="chr1", pos=p, ref=r, alts=a) for p, r, a in variants],
                samples=samples, purities=purities, ploidies=ploidies)


def testx_scenario():
    """REAL alignments: the two BAMs the reference ships (vafator/tests/resources/TESTX_S1_L00{1,2}.bam;
    bwa, 101 bp, proper pairs / orphans / soft clips / indels, MAPQ 0-60), decoded with the product's
    BAM decoder (the container format is not what is under test here).  Read names are replaced by
    surrogates that preserve which records share a name.  Variants: SNVs at covered positions and the
    insertions / deletions the reads themselves carry."""
    from vafator_b200 import bamio
    res = os.path.join(REF_ROOT, "vafator", "tests", "resources")
    bams, contigs, seen = [], None, {}
    for name in ("TESTX_S1_L001.bam", "TESTX_S1_L002.bam"):
        f = bamio.BamFile(os.path.join(res, name))
        contigs = f.contigs
        b = f.read_batch()
        recs = hp.recs_from_soa(b.as_dict())
        out = []
        for r in recs:
            out.append(rec_dict("q%016x" % int(r.name_key[0]), r.flag, r.tid, r.pos, r.mapq,
                                "".join("%d%s" % (l, hp.CIGAR_CHARS[op]) for op, l in r.cigar) or "0M", r.seq, list(r.qual),
                                r.mtid, r.mpos, r.isize))
            if r.flag & 4:
                out[-1]["cigar"] = ""
        bams.append(out)
    rng = random.Random(5)
    variants = {}
    for recs in bams:
        for r in recs:
            if r["flag"] & 4 or not r["cigar"]:
                continue
            ref_pos, q = r["pos"], 0
            ops = hp.parse_cigar(r["cigar"])
            for k, (op, l) in enumerate(ops):
                if op in (0, 7, 8):
                    if rng.random() < 0.12:
                        off = rng.randrange(l)
                        base = r["seq"][q + off]
                        if base in "ACGT":
                            alt = rng.choice([x for x in "ACGT" if x != base])
                            variants.setdefault((r["tid"], ref_pos + off + 1, base), [alt])
                    ref_pos += l
                    q += l
                elif op == 1:
                    if q > 0 and k > 0 and ops[k - 1][0] in (0, 7, 8):
                        anchor = r["seq"][q - 1]
                        variants.setdefault((r["tid"], ref_pos, anchor), [anchor + r["seq"][q:q + l]])
                        variants.setdefault((r["tid"], ref_pos, anchor + "~"), [anchor + r["seq"][q + l:q + 2 * l]])
                    q += l
                elif op == 2:
                    if q > 0 and k > 0 and ops[k - 1][0] in (0, 7, 8):
                        anchor = r["seq"][q - 1]
                        variants.setdefault((r["tid"], ref_pos, anchor + "N" * l), [anchor])
                    ref_pos += l
                elif op == 3:
                    ref_pos += l
                elif op == 4:
                    q += l
    V = []
    for (tid, pos1, ref), alts in variants.items():
        alts = [a for a in alts if len(a) >= 1]
        V.append((tid, pos1, ref.replace("~", ""), alts))
    V = [v for v in V if len(v[2]) >= 1 and all(len(a) > 0 and a != v[2] for a in v[3]) and len(v[3][0]) != 0]
    V = [v for v in V if not (len(v[2]) == 1 and len(v[3][0]) == 1 and v[2] == v[3][0])]
    V.sort(key=lambda v: (v[0], v[1], v[2]))
    return dict(name="testx_real", contigs=contigs, bams=bams,
                variants=[dict(chrom=contigs[t], pos=p, ref=r, alts=a) for t, p, r, a in V],
                samples=OrderedDict([("tumor", [0]), ("normal", [1])]), purities={"tumor": 0.6}, ploidies={"tumor": 3.0})


# ---------------------------------------------------------------------------
# run the reference on a scenario
# ---------------------------------------------------------------------------
def metrics_to_json(m):
    if m is None:
        return None
    f = lambda x: None if (isinstance(x, float) and x != x) else x
    return dict(ac=dict(m.ac), dp=m.dp,
                bqs={k: f(v) for k, v in m.bqs.items()}, mqs={k: f(v) for k, v in m.mqs.items()},
                positions={k: f(v) for k, v in m.positions.items()},
                all_bqs={k: list(map(int, v)) for k, v in m.all_bqs.items()},
                all_mqs={k: list(map(int, v)) for k, v in m.all_mqs.items()},
                all_positions={k: list(map(int, v)) for k, v in m.all_positions.items()})


def run_reference(sc, params, mods):
    rp, ra, rw, rl, rr = mods
    files = [hp.AlignmentFile([to_rec(d, i) for i, d in enumerate(b)], sc["contigs"]) for b in sc["bams"]]
    ploidy_mgrs = {}
    for s, p in sc["ploidies"].items():
        if isinstance(p, list):
            import tempfile
            with tempfile.NamedTemporaryFile("w", suffix=".bed", delete=False) as fh:
                for row in p:
                    fh.write("\t".join(map(str, row)) + "\n")
            ploidy_mgrs[s] = rl.PloidyManager(local_copy_numbers=fh.name)
            os.unlink(fh.name)
        else:
            ploidy_mgrs[s] = rl.PloidyManager(genome_wide_ploidy=p)
    ann = ra.Annotator.__new__(ra.Annotator)
    ann.bam_readers = OrderedDict((s, [files[i] for i in idx]) for s, idx in sc["samples"].items())
    ann.power = rw.PowerCalculator(tumor_ploidies=ploidy_mgrs, purities=dict(sc["purities"]),
                                   normal_ploidy=params.get("normal_ploidy", 2),
                                   fpr=params.get("fpr", rw.DEFAULT_FPR),
                                   error_rate=params.get("error_rate", rw.DEFAULT_ERROR_RATE))
    by_chrom = OrderedDict()
    for v in sc["variants"]:
        by_chrom.setdefault(v["chrom"], []).append(FakeVariant(v["chrom"], v["pos"], v["ref"], v["alts"]))
    per_bam = [dict() for _ in sc["bams"]]
    infos = []
    for chrom, vs in by_chrom.items():
        all_metrics = {}
        for s, idx in sc["samples"].items():
            for j, bi in enumerate(idx):
                res = rp.collect_metrics_for_chrom(chrom=chrom, variants=vs, bam=files[bi],
                                                   min_base_quality=params["min_bq"],
                                                   min_mapping_quality=params["min_mq"],
                                                   include_ambiguous_bases=params["include_ambiguous"])
                all_metrics[(s, j)] = res
                for (pos, ref, alt0), m in res.items():
                    per_bam[bi]["%s:%d:%s:%s" % (chrom, pos, ref, alt0)] = metrics_to_json(m)
        for v in vs:
            mb = {(s, j): all_metrics[(s, j)].get((v.POS, v.REF, v.ALT[0]), ra.EMPTY_METRICS)
                  for s, idx in sc["samples"].items() for j in range(len(idx))}
            ann._add_stats(v, mb)
            infos.append([[k, val] for k, val in v.INFO.items()])
    return dict(params=params, metrics=per_bam, info=infos)


PARAM_SETS = [
    dict(min_bq=30, min_mq=1, include_ambiguous=True),    # CLI defaults (command_line.py:63,71)
    dict(min_bq=29, min_mq=0, include_ambiguous=True),    # Annotator class defaults (annotator.py:76-77)
    dict(min_bq=0, min_mq=0, include_ambiguous=False),    # test_pileups.py:15-16 thresholds
    dict(min_bq=13, min_mq=30, include_ambiguous=False, fpr=1e-5, error_rate=0.01, normal_ploidy=3),
]


def main():
    mods = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    scenarios = [handmade(), random_scenario(11), random_scenario(12, n_bams=2),
                 random_scenario(13, n_bams=3, replicates=True), random_scenario(14, n_pairs=900, read_len=40, G=700)]
    if os.path.exists(os.path.join(REF_ROOT, "vafator", "tests", "resources", "TESTX_S1_L001.bam")):
        scenarios.append(testx_scenario())
    for sc in scenarios:
        sc["runs"] = [run_reference(sc, p, mods) for p in PARAM_SETS]
        path = os.path.join(out_dir, "pileup_%s.json" % sc["name"])
        with open(path, "w") as fh:
            json.dump(sc, fh, separators=(",", ":"))
        print(path, os.path.getsize(path) // 1024, "KiB", len(sc["variants"]), "variants",
              [len(b) for b in sc["bams"]], "reads")
    gen_epilogue_golden(mods, out_dir)


def gen_epilogue_golden(mods, out_dir):
    """Known-answer vectors for the statistics epilogue, straight from the reference modules."""
    import scipy.stats
    rp, ra, rw, rl, rr = mods
    rng = random.Random(20261019)
    rs = []
    for _ in range(400):
        n1, n2 = rng.randint(1, 60), rng.randint(1, 60)
        hi = rng.choice([3, 8, 42, 200])
        alt = [rng.randint(0, hi) for _ in range(n1)]
        ref = [rng.randint(0, hi) for _ in range(n2)]
        z, p = scipy.stats.ranksums(x=alt, y=ref)
        zr, pr = rr.calculate_rank_sum_test(alt, ref)
        rs.append(dict(alt=alt, ref=ref, z=float(z), p=float(p), z_round=float(zr), p_round=float(pr)))
    # the reference's own known answer (vafator/tests/test_rank_sum_test.py:27-35)
    alt = [0, 7, 10, 17, 20, 21, 30, 34, 40, 45]
    ref = [20, 25, 26, 30, 32, 40, 47, 50, 53, 60]
    zr, pr = rr.calculate_rank_sum_test(alt, ref)
    assert (zr, pr) == (-2.154, 0.03121)
    pw = []
    for _ in range(1500):
        purity = rng.choice([0.1, 0.3, 0.5, 0.8, 1.0, rng.random() * 0.99 + 0.01])
        ploidy = rng.choice([0.5, 1, 2, 2.5, 3, 4, 6, rng.random() * 6])
        fpr = rng.choice([5e-7, 1e-5, 1e-3, 1e-9])
        err = rng.choice([1e-3, 3e-3, 0.08, 0.15])
        npl = rng.choice([2, 2, 2, 1, 3])
        dp = rng.choice([0, 1, 2, 5, 10, 30, 60, 100, 500, 1000, 5000, rng.randint(0, 20000)])
        ac = rng.randint(0, dp) if dp else 0
        pc = rw.PowerCalculator(tumor_ploidies={"t": rl.PloidyManager(genome_wide_ploidy=ploidy)},
                                purities={"t": purity}, normal_ploidy=npl, fpr=fpr, error_rate=err)
        eaf = pc.calculate_expected_vaf("t", None)
        pu_raw = float(scipy.stats.binom.cdf(k=ac, n=dp, p=eaf))
        pu = float(pc.calculate_power(dp=dp, ac=ac, sample="t", variant=None))
        if dp > 0:
            k = pc._calculate_k(dp)
            pw_raw = float(1 - scipy.stats.binom.cdf(k=k - 1, n=dp, p=eaf)
                           + pc._calculate_d(k=k, n=dp) * scipy.stats.binom.pmf(k=k, n=dp, p=eaf))
        else:
            pw_raw = None
        pwr, k = pc.calculate_absolute_power("t", None, dp)
        pw.append(dict(purity=purity, ploidy=ploidy, normal_ploidy=npl, fpr=fpr, error_rate=err, dp=dp, ac=ac,
                       eaf=eaf, pu_raw=pu_raw, pu=pu, pw_raw=pw_raw, pw=float(pwr), k=int(k)))
    # a slice of notebooks/power.tsv (reference's published 23 760-row table), every 16th row
    tsv = []
    path = os.path.join(REF_ROOT, "notebooks", "power.tsv")
    if os.path.exists(path):
        with open(path) as fh:
            header = fh.readline().rstrip("\n").split("\t")
            for i, line in enumerate(fh):
                if i % 16 == 0:
                    row = dict(zip(header, line.rstrip("\n").split("\t")))
                    tsv.append({k: float(v) for k, v in row.items() if k in
                                ("purity", "dp", "power", "k", "ploidy", "error_rate")})
    with open(os.path.join(out_dir, "epilogue.json"), "w") as fh:
        json.dump(dict(ranksum=rs, power=pw, power_tsv=tsv), fh, separators=(",", ":"))
    print("epilogue.json", len(rs), len(pw), len(tsv))


if __name__ == "__main__":
    main()
