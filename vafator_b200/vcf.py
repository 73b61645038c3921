"""Minimal text VCF reader / writer (plain, gzip or BGZF) -- stands where cyvcf2's VCF / Writer
stand in the reference (vafator/annotator.py:121-158, :434-441).  Records keep their original
columns; annotations are appended to (or replace keys of) the INFO column, exactly what
htslib's bcf_update_info does for string / integer values."""
from __future__ import annotations

import gzip
from typing import Dict, Iterator, List, Optional


class VcfRecord:
    __slots__ = ("fields", "CHROM", "POS", "REF", "ALT", "INFO")

    def __init__(self, fields: List[str]):
        if len(fields) < 8:
            raise ValueError("malformed VCF record: fewer than 8 columns")
        self.fields = fields
        self.CHROM = fields[0]
        self.POS = int(fields[1])
        self.REF = fields[3]
        self.ALT = [] if fields[4] == "." else fields[4].split(",")
        self.INFO = {}  # annotations to add: {key: value} in insertion order, or finished "key=value;..." text

    def line(self) -> str:
        f = self.fields
        if not self.INFO:
            return "\t".join(f)
        if isinstance(self.INFO, str):  # finished "key=value;..." text (Engine.annotate(as_text=True))
            if f[7] in (".", ""):
                return "\t".join(f[:7] + [self.INFO] + f[8:])
            # appended as it is unless the record already carries one of the keys (then: replaced in place, below);
            # the substring test errs on the safe side only
            if not any((kv.split("=", 1)[0] + "=") in self.INFO for kv in f[7].split(";")):
                return "\t".join(f[:7] + [f[7] + ";" + self.INFO] + f[8:])
            self.INFO = dict(kv.split("=", 1) if "=" in kv else (kv, "") for kv in self.INFO.split(";"))
        if f[7] in (".", ""):
            info = ";".join(["%s=%s" % kv for kv in self.INFO.items()])
        else:
            old = f[7].split(";")
            keys = {kv.split("=", 1)[0]: i for i, kv in enumerate(old)}
            if keys.keys().isdisjoint(self.INFO):
                info = f[7] + ";" + ";".join(["%s=%s" % kv for kv in self.INFO.items()])
            else:  # an annotation replaces the value of a key the record already carries, in place
                for k, v in self.INFO.items():
                    kv = "%s=%s" % (k, v)
                    if k in keys:
                        old[keys[k]] = kv
                    else:
                        old.append(kv)
                info = ";".join(old)
        return "\t".join(f[:7] + [info] + f[8:])


class VcfReader:
    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as fh:
            magic = fh.read(2)
        self._fh = gzip.open(path, "rt") if magic == b"\x1f\x8b" else open(path, "rt")
        self.header_lines: List[str] = []
        self.column_line: Optional[str] = None
        for line in self._fh:
            line = line.rstrip("\n").rstrip("\r")
            if line.startswith("##"):
                self.header_lines.append(line)
            elif line.startswith("#"):
                self.column_line = line
                break
            else:
                raise ValueError("%s: VCF record before the #CHROM header line" % path)
        if self.column_line is None:
            raise ValueError("%s: no #CHROM header line" % path)
        self.extra_header: List[str] = []

    def add_to_header(self, line: str) -> None:
        self.extra_header.append(line)

    def add_info_to_header(self, d: dict) -> None:
        self.extra_header.append('##INFO=<ID={ID},Number={Number},Type={Type},Description="{Description}">'.format(**d))

    def __iter__(self) -> Iterator[VcfRecord]:
        for line in self._fh:
            line = line.rstrip("\n").rstrip("\r")
            if line:
                yield VcfRecord(line.split("\t"))

    def close(self) -> None:
        self._fh.close()


class VcfWriter:
    def __init__(self, path: str, template: VcfReader):
        self.path = path
        self._bgzf = path.endswith(".gz")
        self._raw = open(path, "wb" if self._bgzf else "w")
        self._pending = bytearray()  # BGZF: bytes not yet written as a block
        self._emit("\n".join(template.header_lines + template.extra_header + [template.column_line]) + "\n")

    def _emit(self, text: str) -> None:
        if not self._bgzf:
            self._raw.write(text)
            return
        from .bamio import _bgzf_block
        self._pending += text.encode()
        while len(self._pending) >= 0xFF00:
            self._raw.write(_bgzf_block(bytes(self._pending[:0xFF00])))
            del self._pending[:0xFF00]

    def write_record(self, rec: VcfRecord) -> None:
        self._emit(rec.line() + "\n")

    def close(self) -> None:
        if self._bgzf:
            from .bamio import _bgzf_block, _BGZF_EOF
            if self._pending:
                self._raw.write(_bgzf_block(bytes(self._pending)))
                self._pending.clear()
            self._raw.write(_BGZF_EOF)
        self._raw.close()
