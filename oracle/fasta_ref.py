"""Independent (pure Python/numpy) FASTA reader used only to check the product parser.
TEST INFRASTRUCTURE ONLY.

Restates what the reference gets from seq_io (src/main.rs:62-72, src/protein.rs:107-110,
135-138): id = header up to the first whitespace; class = 4th '|' field of the id
(`split_terminator('|')[3]`); sequence = the record's residue bytes (line breaks removed,
SURVEY C4).  Class ids are assigned in first-occurrence order.
"""
from __future__ import annotations

import numpy as np


def parse_fasta_bytes(data: bytes):
    ids, seqs, cur = [], [], None
    for line in data.split(b"\n"):
        line = line.rstrip(b"\r")
        if line.startswith(b">"):
            if cur is not None:
                seqs.append(b"".join(cur))
            hdr = line[1:]
            ids.append(hdr.split(None, 1)[0].decode() if hdr.strip() else "")
            cur = []
        elif cur is not None and line:
            cur.append(line)
    if cur is not None:
        seqs.append(b"".join(cur))
    classes, class_names, table = [], [], {}
    for pid in ids:
        f = pid.split("|")
        if f and f[-1] == "":      # split_terminator drops one trailing empty field
            f = f[:-1]
        name = f[3] if len(f) > 3 else ""
        if name not in table:
            table[name] = len(class_names)
            class_names.append(name)
        classes.append(table[name])
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint64)
    np.cumsum(lens, out=offsets[1:])
    residues = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy()
    return {"ids": ids, "residues": residues, "offsets": offsets,
            "class_id": np.array(classes, dtype=np.uint32), "class_names": class_names}


def parse_fasta(path: str):
    with open(path, "rb") as fh:
        return parse_fasta_bytes(fh.read())
