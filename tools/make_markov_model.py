"""Trains the order-2 Markov model of SURVEY 8(d).2 on the reference's four English corpus files and
stores it as zzflate_b200/data/markov2.npz (derived statistics only -- the corpus text is not stored).
Run in the dev container (needs /root/reference)."""
import sys
from pathlib import Path
import numpy as np

root = Path(__file__).resolve().parent.parent
corpus = Path("/root/reference/zztest/corpus")
text = b"".join((corpus / f).read_bytes() for f in ("alice29.txt", "asyoulik.txt", "lcet10.txt", "plrabn12.txt"))
a = np.frombuffer(text + text[:2], dtype=np.uint8).astype(np.uint32)     # circular: every context has a successor
ctx = (a[:-2] << 8) | a[1:-1]
key = (ctx << 8) | a[2:]
uniq, counts = np.unique(key, return_counts=True)
rows = uniq >> 8
syms = (uniq & 0xFF).astype(np.uint8)
row_off = np.zeros(65537, dtype=np.uint32)
np.add.at(row_off, rows + 1, 1)
row_off = np.cumsum(row_off).astype(np.uint32)
np.savez_compressed(root / "zzflate_b200" / "data" / "markov2.npz", row_off=row_off, syms=syms,
                    counts=counts.astype(np.uint32), start=np.frombuffer(text[:2], dtype=np.uint8))
print("contexts", len(np.unique(rows)), "entries", len(uniq), "training bytes", len(text))
