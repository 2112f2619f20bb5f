"""Marker code tables (host side, numpy only).

Mirrors `ARDictionary` of the reference (`/root/reference/src/dictionaries.rs:22-28`, map at
`:30-113`, constructor `:116-145`, `get_mark_size` `:154-156`).  The tables themselves are data
extracted by `tools/extract_dictionaries.py` into `data/dictionaries.bin`; the very same blob is
embedded into the CUDA library (`csrc/a3_dictionary.cpp`), so both sides agree by construction.
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass
from functools import lru_cache
from pathlib import Path

import numpy as np

_BLOB = Path(__file__).resolve().parent / "data" / "dictionaries.bin"
_ENTRY = struct.Struct("<24sBBHIII")


@dataclass(frozen=True)
class DictionaryTable:
    name: str
    num_bits: int
    tau_table: int          # 0 = "compute as the minimum pairwise distance" (dictionaries.rs:124)
    codes: np.ndarray       # uint64[n]

    @property
    def mark_size(self) -> int:
        """`(num_bits as f32).sqrt().ceil() as u8 + 2` (dictionaries.rs:154-156)."""
        return int(math.ceil(math.sqrt(np.float32(self.num_bits)))) + 2

    @property
    def tau(self) -> int:
        return self.tau_table if self.tau_table else calculate_tau(self.codes)


@lru_cache(maxsize=None)
def _load() -> dict:
    blob = _BLOB.read_bytes()
    if blob[:8] != b"A3DICT01":
        raise RuntimeError(f"{_BLOB}: bad magic")
    n_entries, n_codes = struct.unpack_from("<II", blob, 8)
    codes = np.frombuffer(blob, dtype="<u8", count=n_codes, offset=16 + _ENTRY.size * n_entries)
    out = {}
    for i in range(n_entries):
        name, nb, tau, _, n, first, _ = _ENTRY.unpack_from(blob, 16 + _ENTRY.size * i)
        name = name.rstrip(b"\0").decode()
        out[name] = DictionaryTable(name, nb, tau, codes[first:first + n])
    return out


def dictionary_names() -> list:
    return sorted(_load())


def table(name: str) -> DictionaryTable:
    """Case-insensitive lookup like `new_from_named_dict` (dictionaries.rs:140-145)."""
    t = _load().get(name.upper())
    if t is None:
        raise KeyError(f"unknown dictionary {name!r}")
    return t


@lru_cache(maxsize=None)
def _tau_cached(key: bytes) -> int:
    codes = np.frombuffer(key, dtype="<u8")
    best = 255
    for i in range(len(codes) - 1):
        d = int(np.bitwise_count(codes[i + 1:] ^ codes[i]).min())
        best = min(best, d)
    return best


def calculate_tau(codes: np.ndarray) -> int:
    """Minimum pairwise Hamming distance (dictionaries.rs:129-138)."""
    return _tau_cached(np.ascontiguousarray(codes, dtype="<u8").tobytes())
