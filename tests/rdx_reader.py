"""Minimal reader of R's RDX2 (XDR) serialisation — just enough for data/germancredit.RData of the reference (a pairlist of
an integer vector and a numeric matrix with attributes).  Test infrastructure."""
import gzip
import struct

import numpy as np


class RDX:
    def __init__(self, path):
        self.b, self.o, self.refs = gzip.decompress(open(path, "rb").read()), 0, []
        assert self.b[:7] == b"RDX2\nX\n"
        self.o = 7 + 12

    def i32(self):
        v = struct.unpack(">i", self.b[self.o:self.o + 4])[0]
        self.o += 4
        return v

    def item(self):
        flags = self.i32()
        t, has_attr, has_tag = flags & 0xff, bool(flags & 0x200), bool(flags & 0x400)
        if t == 254:
            return None
        if t == 255:
            return self.refs[(flags >> 8) - 1]
        if t == 1:
            s = self.item()
            self.refs.append(s)
            return s
        if t == 9:
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.o:self.o + n].decode("latin1")
            self.o += n
            return s
        if t == 2:
            out = {}
            while True:
                if has_attr:
                    self.item()
                tag = self.item() if has_tag else None
                out[tag] = self.item()
                flags = self.i32()
                if flags & 0xff == 254:
                    return out
                assert flags & 0xff == 2
                has_attr, has_tag = bool(flags & 0x200), bool(flags & 0x400)
        n = self.i32()
        if t in (10, 13):
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.o).astype(np.int64)
            self.o += 4 * n
        elif t == 14:
            v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.o).astype(np.float64)
            self.o += 8 * n
        elif t in (16, 19):
            v = [self.item() for _ in range(n)]
        else:
            raise ValueError(f"unsupported SEXP type {t}")
        attrs = self.item() if has_attr else None
        return (v, attrs) if attrs else v


def read_rdata(path):
    return RDX(path).item()
