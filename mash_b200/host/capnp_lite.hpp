// capnp_lite.hpp -- a minimal Cap'n Proto encoder/decoder, just enough for Mash's MinHash schema
// (reference src/mash/capnp/MinHash.capnp, file id 0xc4c8b1ada05e7704), because libcapnp/libkj are not available.
//
// Wire format (public Cap'n Proto encoding spec): 64-bit words; struct pointer = {offset:30 (signed, words, from the
// end of the pointer), kind 0, dataWords:16, ptrCount:16}; list pointer = {offset:30, kind 1, elemSize:3, count:29};
// far pointer = {kind 2, doubleFar:1, padOffset:29, segmentId:32}; composite lists start with a tag word shaped like
// a struct pointer whose offset field is the element count; Text = byte list including the NUL; stream framing =
// u32 (segments-1), u32 size[segments], pad to 8 bytes, segment payloads.
//
// Builder: mimics capnp::MallocMessageBuilder + WireHelpers::allocate so that files come out segment-for-segment like
// the reference writer's (Sketch.cpp:384-490): first segment 1024 words, later segments max(needed, words allocated
// so far); an object is first tried in the segment of the pointer that refers to it, then in the newest segment, else
// a new segment -- in the last two cases behind a far pointer + landing pad.  (Derived from the documented behaviour;
// no reference-written .msh exists in the tree to diff against, SURVEY.md 5.1 / appendix B.)
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace capnp_lite {

typedef uint64_t word;

// ------------------------------------------------------------------------------------------------ builder
class Builder {
public:
    struct Ptr { uint32_t seg; uint32_t off; };   // location of a word
    static constexpr uint32_t FIRST_SEGMENT_WORDS = 1024;

    Builder() : nextSize_(FIRST_SEGMENT_WORDS) {}

    // allocates the root pointer word (segment 0, word 0)
    Ptr initRootPointer() { return arenaAllocate(1); }

    word &at(Ptr p) { return segs_[p.seg].w[p.off]; }
    Ptr plus(Ptr p, uint32_t n) const { return Ptr{p.seg, p.off + n}; }

    // allocate a struct for the pointer at `ref`; returns location of its first data word
    Ptr initStruct(Ptr ref, uint16_t dataWords, uint16_t ptrCount)
    {
        Ptr target = allocateFor(ref, (uint32_t)dataWords + ptrCount, [&](int64_t off) {
            return (uint64_t)((uint32_t)(off << 2) | 0u) | ((uint64_t)dataWords << 32) | ((uint64_t)ptrCount << 48);
        });
        return target;
    }

    // list of primitive elements; elemSizeCode: 2 = byte, 4 = four bytes, 5 = eight bytes
    Ptr initList(Ptr ref, int elemSizeCode, uint32_t count)
    {
        static const int bits[8] = {0, 1, 8, 16, 32, 64, 64, 0};
        uint64_t totalBits = (uint64_t)bits[elemSizeCode] * count;
        uint32_t words = (uint32_t)((totalBits + 63) / 64);
        return allocateFor(ref, words, [&](int64_t off) {
            return (uint64_t)((uint32_t)(off << 2) | 1u) | ((uint64_t)elemSizeCode << 32) | ((uint64_t)count << 35);
        });
    }

    // composite (struct) list; returns location of the first element's first word (after the tag)
    Ptr initStructList(Ptr ref, uint32_t count, uint16_t dataWords, uint16_t ptrCount)
    {
        uint32_t per = (uint32_t)dataWords + ptrCount;
        uint32_t words = count * per;
        Ptr tag = allocateFor(ref, words + 1, [&](int64_t off) {
            return (uint64_t)((uint32_t)(off << 2) | 1u) | ((uint64_t)7 << 32) | ((uint64_t)words << 35);
        });
        at(tag) = (uint64_t)((uint32_t)(count << 2) | 0u) | ((uint64_t)dataWords << 32) | ((uint64_t)ptrCount << 48);
        return plus(tag, 1);
    }

    void setText(Ptr ref, const std::string &s)
    {
        Ptr p = initList(ref, 2, (uint32_t)s.size() + 1);
        memcpy(reinterpret_cast<char *>(&at(p)), s.data(), s.size());   // NUL already there (zeroed segment)
    }

    // serialise: framing table + used words of each segment
    std::string serialize() const
    {
        std::string out;
        uint32_t n = (uint32_t)segs_.size();
        std::vector<uint32_t> table;
        table.push_back(n - 1);
        for (auto &s : segs_) table.push_back(s.used);
        if (table.size() % 2) table.push_back(0);
        out.append(reinterpret_cast<const char *>(table.data()), table.size() * 4);
        for (auto &s : segs_) out.append(reinterpret_cast<const char *>(s.w.data()), (size_t)s.used * 8);
        return out;
    }

    size_t segmentCount() const { return segs_.size(); }
    uint32_t segmentUsed(size_t i) const { return segs_[i].used; }

private:
    struct Segment { std::vector<word> w; uint32_t used = 0; };
    std::vector<Segment> segs_;
    uint32_t nextSize_;

    bool tryAlloc(uint32_t seg, uint32_t amount, uint32_t &off)
    {
        Segment &s = segs_[seg];
        if (s.w.size() - s.used < amount) return false;
        off = s.used;
        s.used += amount;
        return true;
    }

    // MallocMessageBuilder::allocateSegment with GROW_HEURISTICALLY
    void addSegment(uint32_t minimum)
    {
        uint32_t size = minimum > nextSize_ ? minimum : nextSize_;
        Segment s;
        s.w.assign(size, 0);
        if (segs_.empty()) nextSize_ = size; else nextSize_ += size;
        segs_.push_back(std::move(s));
    }

    // BuilderArena::allocate: only the newest segment is tried before a new one is made
    Ptr arenaAllocate(uint32_t amount)
    {
        uint32_t off = 0;
        if (segs_.empty()) addSegment(amount);
        else if (!tryAlloc((uint32_t)segs_.size() - 1, amount, off)) { addSegment(amount); }
        else return Ptr{(uint32_t)segs_.size() - 1, off};
        tryAlloc((uint32_t)segs_.size() - 1, amount, off);
        return Ptr{(uint32_t)segs_.size() - 1, off};
    }

    // WireHelpers::allocate
    template <typename MakePtr>
    Ptr allocateFor(Ptr ref, uint32_t amount, MakePtr make)
    {
        uint32_t off;
        if (tryAlloc(ref.seg, amount, off)) {
            at(ref) = make((int64_t)off - ((int64_t)ref.off + 1));
            return Ptr{ref.seg, off};
        }
        Ptr pad = arenaAllocate(amount + 1);
        at(ref) = (uint64_t)(((uint32_t)pad.off << 3) | 2u) | ((uint64_t)pad.seg << 32);   // far pointer, single
        at(pad) = make(0);                                                                 // landing pad: object follows it
        return Ptr{pad.seg, pad.off + 1};
    }
};

// ------------------------------------------------------------------------------------------------ reader
class Reader {
public:
    struct StructRef { const Reader *r = nullptr; uint32_t seg = 0; uint32_t off = 0; uint16_t dataWords = 0, ptrCount = 0; bool valid = false; };
    struct ListRef { const Reader *r = nullptr; uint32_t seg = 0; uint32_t off = 0; int elemSize = 0; uint32_t count = 0; uint16_t dataWords = 0, ptrCount = 0; bool valid = false; };

    Reader(const void *data, size_t bytes)
    {
        const uint8_t *p = static_cast<const uint8_t *>(data);
        if (bytes < 8) throw std::runtime_error("message too short");
        uint32_t nseg;
        memcpy(&nseg, p, 4);
        nseg += 1;
        size_t table = 4 + 4ull * nseg;
        table = (table + 7) & ~size_t(7);
        if (nseg > 1u << 20 || bytes < table) throw std::runtime_error("bad segment table");
        size_t pos = table;
        for (uint32_t i = 0; i < nseg; i++) {
            uint32_t sz;
            memcpy(&sz, p + 4 + 4ull * i, 4);
            if (pos + 8ull * sz > bytes) throw std::runtime_error("segment exceeds message");
            segs_.push_back({reinterpret_cast<const word *>(p + pos), sz});
            pos += 8ull * sz;
        }
    }

    StructRef root() const { return followStruct(0, 0); }

    // data fields (little endian); fields beyond the struct's data section read as 0 (schema evolution rule)
    template <typename T> T get(const StructRef &s, uint32_t byteOffset) const
    {
        T v = 0;
        if (!s.valid || byteOffset + sizeof(T) > 8u * s.dataWords) return v;
        memcpy(&v, reinterpret_cast<const uint8_t *>(segs_[s.seg].w + s.off) + byteOffset, sizeof(T));
        return v;
    }
    bool getBit(const StructRef &s, uint32_t bit) const { return (get<uint8_t>(s, bit / 8) >> (bit % 8)) & 1; }

    bool pointerIsNull(const StructRef &s, uint32_t index) const
    {
        if (!s.valid || index >= s.ptrCount) return true;
        return segs_[s.seg].w[s.off + s.dataWords + index] == 0;
    }
    StructRef getStruct(const StructRef &s, uint32_t index) const
    {
        if (pointerIsNull(s, index)) return StructRef();
        return followStruct(s.seg, s.off + s.dataWords + index);
    }
    ListRef getList(const StructRef &s, uint32_t index) const
    {
        if (pointerIsNull(s, index)) return ListRef();
        return followList(s.seg, s.off + s.dataWords + index);
    }
    std::string getText(const StructRef &s, uint32_t index) const
    {
        ListRef l = getList(s, index);
        if (!l.valid || l.elemSize != 2 || l.count == 0) return std::string();
        const char *c = reinterpret_cast<const char *>(segs_[l.seg].w + l.off);
        return std::string(c, l.count - 1);
    }
    StructRef element(const ListRef &l, uint32_t i) const
    {
        StructRef s;
        if (!l.valid || l.elemSize != 7 || i >= l.count) return s;
        s.r = this; s.seg = l.seg; s.off = l.off + i * ((uint32_t)l.dataWords + l.ptrCount);
        s.dataWords = l.dataWords; s.ptrCount = l.ptrCount; s.valid = true;
        return s;
    }
    // primitive list elements; the element size code of the list must be the one asked for (5 = 64 bit, 4 = 32 bit) and the
    // index inside the list (followList checked that `count` elements lie inside the segment) -- as libcapnp bounds-checks
    uint64_t elementU64(const ListRef &l, uint32_t i) const
    {
        if (!l.valid || l.elemSize != 5 || i >= l.count) throw std::runtime_error("bad 64-bit list access");
        uint64_t v; memcpy(&v, reinterpret_cast<const uint8_t *>(segs_[l.seg].w + l.off) + 8ull * i, 8); return v;
    }
    uint32_t elementU32(const ListRef &l, uint32_t i) const
    {
        if (!l.valid || l.elemSize != 4 || i >= l.count) throw std::runtime_error("bad 32-bit list access");
        uint32_t v; memcpy(&v, reinterpret_cast<const uint8_t *>(segs_[l.seg].w + l.off) + 4ull * i, 4); return v;
    }

private:
    struct Seg { const word *w; uint32_t n; };
    std::vector<Seg> segs_;

    // resolve far pointers; returns the effective pointer word and the location its offset is relative to
    void resolve(uint32_t seg, uint32_t off, uint64_t &ptr, uint32_t &tseg, int64_t &base) const
    {
        if (seg >= segs_.size() || off >= segs_[seg].n) throw std::runtime_error("pointer out of bounds");
        ptr = segs_[seg].w[off];
        tseg = seg;
        base = (int64_t)off + 1;
        if ((ptr & 3) != 2) return;
        bool dbl = (ptr >> 2) & 1;
        uint32_t padOff = (uint32_t)(ptr & 0xFFFFFFFFu) >> 3;
        uint32_t padSeg = (uint32_t)(ptr >> 32);
        if (padSeg >= segs_.size() || padOff + (dbl ? 1u : 0u) >= segs_[padSeg].n) throw std::runtime_error("far pointer out of bounds");
        if (!dbl) {
            ptr = segs_[padSeg].w[padOff];
            tseg = padSeg;
            base = (int64_t)padOff + 1;
        } else {
            uint64_t far2 = segs_[padSeg].w[padOff];
            ptr = segs_[padSeg].w[padOff + 1];          // tag: offset field unused
            tseg = (uint32_t)(far2 >> 32);
            base = (int64_t)((uint32_t)(far2 & 0xFFFFFFFFu) >> 3);
            ptr &= ~0xFFFFFFFCull;                        // zero offset: object starts at `base`
        }
    }

    StructRef followStruct(uint32_t seg, uint32_t off) const
    {
        uint64_t ptr; uint32_t tseg; int64_t base;
        resolve(seg, off, ptr, tseg, base);
        StructRef s;
        if (ptr == 0) return s;
        if ((ptr & 3) != 0) throw std::runtime_error("expected struct pointer");
        int32_t o = (int32_t)(uint32_t)(ptr & 0xFFFFFFFFu) >> 2;
        s.r = this; s.seg = tseg; s.off = (uint32_t)(base + o);
        s.dataWords = (uint16_t)(ptr >> 32); s.ptrCount = (uint16_t)(ptr >> 48);
        if ((uint64_t)s.off + s.dataWords + s.ptrCount > segs_[tseg].n) throw std::runtime_error("struct out of bounds");
        s.valid = true;
        return s;
    }

    ListRef followList(uint32_t seg, uint32_t off) const
    {
        uint64_t ptr; uint32_t tseg; int64_t base;
        resolve(seg, off, ptr, tseg, base);
        ListRef l;
        if (ptr == 0) return l;
        if ((ptr & 3) != 1) throw std::runtime_error("expected list pointer");
        int32_t o = (int32_t)(uint32_t)(ptr & 0xFFFFFFFFu) >> 2;
        l.r = this; l.seg = tseg; l.off = (uint32_t)(base + o);
        l.elemSize = (int)((ptr >> 32) & 7);
        uint32_t count = (uint32_t)(ptr >> 35);
        if (l.elemSize == 7) {
            if (l.off >= segs_[tseg].n) throw std::runtime_error("list out of bounds");
            uint64_t tag = segs_[tseg].w[l.off];
            l.count = (uint32_t)(tag & 0xFFFFFFFFu) >> 2;
            l.dataWords = (uint16_t)(tag >> 32); l.ptrCount = (uint16_t)(tag >> 48);
            if ((uint64_t)l.off + 1 + count > segs_[tseg].n) throw std::runtime_error("list out of bounds");
            // `count` of a composite list is its size in words: the tag's element count times the element size must fit in it
            if ((uint64_t)l.count * ((uint64_t)l.dataWords + l.ptrCount) > count) throw std::runtime_error("composite list larger than its word count");
            l.off += 1;
        } else {
            static const int bits[8] = {0, 1, 8, 16, 32, 64, 64, 0};
            l.count = count;
            uint64_t words = ((uint64_t)bits[l.elemSize] * count + 63) / 64;
            if ((uint64_t)l.off + words > segs_[tseg].n) throw std::runtime_error("list out of bounds");
        }
        l.valid = true;
        return l;
    }
};

}  // namespace capnp_lite
