"""INFO contract of the annotation (mirrors vafator/constants.py:2-115: same IDs, Number, Type
and Description text, because downstream tools parse these header lines)."""

BATCH_SIZE = 10000  # records buffered between flushes of the output VCF (vafator/constants.py:2)

# IUPAC ambiguity codes counted in `{sample}_n` (vafator/constants.py:5)
AMBIGUOUS_BASES = ["N", "M", "R", "W", "S", "Y", "K", "V", "H", "D", "B"]

_S = "in the {sample} sample/s"
_RS = "Rank sum test comparing the %s distributions supporting the reference and the alternate " + _S
_RSPV = ("Rank sum test p-value for %s distributions " + _S +
         ". The null hypothesis is that there is no difference between the distributions")

# (suffix, description, type, number) in the reference's order
_HEADER_TEMPLATES = [
    ("af", "Allele frequency for the alternate alleles " + _S, "Float", "A"),
    ("dp", "Total depth of coverage " + _S + " (independent of alleles)", "Float", "1"),
    ("ac", "Allele count for the alternate alleles " + _S, "Integer", "A"),
    ("n", "Allele count for ambiguous bases (any IUPAC ambiguity code is counted) " + _S, "Integer", "1"),
    ("pu", "Probability of an undetected mutation given the observed supporting reads (AC), the observed total "
           "coverage (DP) and the provided tumor purity " + _S, "Float", "A"),
    ("pw", "Power to detect a somatic mutation as described in Absolute given the observed total coverage (DP) "
           "and the provided tumor purity and ploidies " + _S, "Float", "1"),
    ("k", "Minimum number of supporting reads, k, such that the probability of observing k or more "
          "non-reference reads due to sequencing error is less than the defined FPR " + _S, "Float", "1"),
    ("eaf", "Expected VAF considering the purity and ploidy/copy number " + _S, "Float", "1"),
    ("bq", "Median base call quality of the reads supporting each allele " + _S, "Float", "R"),
    ("mq", "Median mapping quality of the reads supporting each allele " + _S, "Float", "R"),
    ("pos", "Median position within the read of the reads supporting each allele " + _S, "Float", "R"),
    ("rsmq", _RS % "MQ", "Float", "A"),
    ("rsmq_pv", _RSPV % "MQ", "Float", "A"),
    ("rsbq", _RS % "BQ", "Float", "A"),
    ("rsbq_pv", _RSPV % "BQ", "Float", "A"),
    ("rspos", _RS % "position", "Float", "A"),
    ("rspos_pv", _RSPV % "position", "Float", "A"),
]

# eaf is not produced per replicate (vafator/constants.py:115)
_REPLICATE_HEADER_TEMPLATES = [t for t in _HEADER_TEMPLATES if t[0] != "eaf"]
