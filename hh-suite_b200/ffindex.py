"""Reader for ffindex databases (`X_hhm.ffdata` + `X_hhm.ffindex`, `X_cs219.ff*`), the container format every
HH-suite database ships in (lib/ffindex/src/ffindex.h: ffindex_entry_t {offset, length, name}; the index is a
text file of `name\\toffset\\tlength` lines, lengths include the trailing NUL byte of each record)."""
from __future__ import annotations

import mmap
import os

import numpy as np


class FFIndex:
    def __init__(self, data_path: str, index_path: str | None = None):
        index_path = index_path or os.path.splitext(data_path)[0] + ".ffindex"
        names, off, ln = [], [], []
        with open(index_path, "rb") as f:
            for line in f:
                parts = line.rstrip(b"\n").split(b"\t")
                if len(parts) != 3:
                    raise ValueError(f"{index_path}: malformed index line {line!r}")
                names.append(parts[0].decode())
                off.append(int(parts[1])); ln.append(int(parts[2]))
        self.names = names
        self.offsets = np.array(off, np.int64)
        self.lengths = np.array(ln, np.int64)
        self._f = open(data_path, "rb")
        size = os.fstat(self._f.fileno()).st_size
        self.data = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ) if size else b""
        if len(self.names) and int((self.offsets + self.lengths).max()) > size:
            raise ValueError(f"{index_path}: an entry ends beyond {data_path}")

    def __len__(self):
        return len(self.names)

    def record(self, k: int) -> bytes:
        return self.data[self.offsets[k]:self.offsets[k] + self.lengths[k]]

    def close(self):
        if hasattr(self.data, "close"):
            try:
                self.data.close()
            except BufferError:       # numpy views of the mapping are still alive; the map goes with the last of them
                pass
        self._f.close()


def write_ffindex(data_path: str, records: list[tuple[str, bytes]], index_path: str | None = None):
    """Write records as an ffindex pair (entries sorted by name, NUL-terminated like ffindex_build)."""
    index_path = index_path or os.path.splitext(data_path)[0] + ".ffindex"
    entries = []
    with open(data_path, "wb") as d:
        pos = 0
        for name, blob in records:
            d.write(blob + b"\0")
            entries.append((name, pos, len(blob) + 1))
            pos += len(blob) + 1
    with open(index_path, "w") as f:
        for name, off, ln in sorted(entries):
            f.write(f"{name}\t{off}\t{ln}\n")


def cs219_arrays(ff: FFIndex):
    """(L, off, seq) for capi.CsDB from an `X_cs219` ffindex in the binary column-state format, as
    Prefilter::init_prefilter reads it (src/hhprefilter.cpp:314-335): record n = `length[n] = entry->length - 1`
    state bytes followed by NUL.  The old text format (records starting with '>') is refused like checkCSFormat does
    (:337-352)."""
    seq = np.frombuffer(ff.data, np.uint8)
    L = (ff.lengths - 1).astype(np.int32)
    if len(L) and L.min() < 1:
        raise ValueError("cs219 database has an empty record")
    for k in range(min(5, len(ff))):
        if seq[ff.offsets[k]] == ord(">"):
            raise ValueError("cs219 database is in the old text format; rebuild it with cstranslate -b (binary)")
    return L, ff.offsets.copy(), seq


def hhm_arrays(ff: FFIndex):
    """(data, offsets, lengths) of an `X_hhm` ffindex for capi.TargetDB.from_hhm."""
    return ff.data, ff.offsets.copy(), ff.lengths.copy()
