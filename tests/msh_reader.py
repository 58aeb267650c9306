"""Independent (pure Python) decoder of Mash .msh files, written from the public Cap'n Proto encoding spec and the
reference schema (src/mash/capnp/MinHash.capnp).  Used to cross-check the C++ writer in mash_b200/host."""
import struct


class Msg:
    def __init__(self, data):
        n = struct.unpack_from("<I", data, 0)[0] + 1
        sizes = struct.unpack_from("<%dI" % n, data, 4)
        pos = (4 + 4 * n + 7) & ~7
        self.segs = []
        for s in sizes:
            self.segs.append(data[pos:pos + 8 * s])
            pos += 8 * s
        self.sizes = sizes
        assert pos == len(data), "trailing bytes"

    def word(self, seg, off):
        return struct.unpack_from("<Q", self.segs[seg], 8 * off)[0]

    def resolve(self, seg, off):
        p = self.word(seg, off)
        base = off + 1
        if p & 3 == 2:
            assert not (p >> 2) & 1, "double-far not expected from this writer"
            pad_off, pad_seg = (p & 0xFFFFFFFF) >> 3, p >> 32
            p = self.word(pad_seg, pad_off)
            seg, base = pad_seg, pad_off + 1
        return p, seg, base

    @staticmethod
    def _soff(p):
        o = (p & 0xFFFFFFFF) >> 2
        return o - (1 << 30) if o & (1 << 29) else o

    def struct(self, seg, off):
        p, seg, base = self.resolve(seg, off)
        if p == 0:
            return None
        assert p & 3 == 0
        return dict(seg=seg, off=base + self._soff(p), dw=(p >> 32) & 0xFFFF, pc=p >> 48)

    def list(self, seg, off):
        p, seg, base = self.resolve(seg, off)
        if p == 0:
            return None
        assert p & 3 == 1
        start = base + self._soff(p)
        es, cnt = (p >> 32) & 7, p >> 35
        if es == 7:
            tag = self.word(seg, start)
            return dict(seg=seg, off=start + 1, es=7, count=(tag & 0xFFFFFFFF) >> 2, dw=(tag >> 32) & 0xFFFF, pc=tag >> 48)
        return dict(seg=seg, off=start, es=es, count=cnt)

    def text(self, seg, off):
        l = self.list(seg, off)
        if l is None:
            return None
        assert l["es"] == 2
        raw = self.segs[l["seg"]][8 * l["off"]:8 * l["off"] + l["count"]]
        assert raw[-1] == 0
        return raw[:-1].decode()

    def prim_list(self, seg, off):
        l = self.list(seg, off)
        if l is None:
            return None
        fmt = {4: "I", 5: "Q"}[l["es"]]
        return list(struct.unpack_from("<%d%s" % (l["count"], fmt), self.segs[l["seg"]], 8 * l["off"]))


def read_msh(path):
    m = Msg(open(path, "rb").read())
    root = m.struct(0, 0)
    assert (root["dw"], root["pc"]) == (3, 4)
    d = m.segs[root["seg"]]
    b = 8 * root["off"]
    kmer, window, mhpw, flags, error_bits, seed = struct.unpack_from("<IIIIII", d, b)
    p0 = root["off"] + root["dw"]
    out = dict(kmer=kmer, windowSize=window, sketchSize=mhpw, concatenated=bool(flags & 1), noncanonical=bool(flags & 2),
               preserveCase=bool(flags & 4), hashSeed=seed ^ 42, alphabet=m.text(root["seg"], p0 + 2), segments=list(m.sizes))
    which = None
    refs = []
    for slot, name in ((3, "referenceList"), (0, "referenceListOld")):
        rl = m.struct(root["seg"], p0 + slot)
        if rl is None:
            continue
        lst = m.list(rl["seg"], rl["off"] + rl["dw"])
        if lst is not None and lst["count"]:
            which = name
            assert (lst["dw"], lst["pc"]) == (2, 7)
            for i in range(lst["count"]):
                o = lst["off"] + 9 * i
                length32, bits = struct.unpack_from("<II", m.segs[lst["seg"]], 8 * o)
                length64 = struct.unpack_from("<Q", m.segs[lst["seg"]], 8 * (o + 1))[0]
                po = o + 2
                refs.append(dict(name=m.text(lst["seg"], po + 2), comment=m.text(lst["seg"], po + 3), length=length64 or length32,
                                 hashes32=m.prim_list(lst["seg"], po + 4), hashes64=m.prim_list(lst["seg"], po + 5),
                                 counts32=m.prim_list(lst["seg"], po + 6), counts32Sorted=bool(bits & 1)))
            break
    out["list"] = which
    out["references"] = refs
    ll = m.struct(root["seg"], p0 + 1)
    out["hasLocusList"] = ll is not None
    return out
