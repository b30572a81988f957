// tiff_io.hpp -- SURVEY 8(f) N3: the TIFF files of the reference's task flow, without libtiff / GDAL.
//
// Writes what cv::imwrite (ref preproc.h:167-185, imageop.h:444) and the GTiff driver (ref imageop.h:316-328, :470-538)
// are used for there: 16-bit unsigned rasters with 1 or 4 interleaved samples per pixel, as baseline TIFF strips
// (classic TIFF below 4 GiB, BigTIFF above), uncompressed (the single-band GTiff of ref imageop.h:316-328) or with the
// options the reference's libraries apply to the other products: COMPRESS=LZW + PREDICTOR=2 (ref imageop.h:470-474;
// cv::imwrite's TIFF default).  The LZW coder follows libtiff's (MSB-first 9..12-bit codes, early change, one stream per
// strip); strips are compressed on all host cores.  Files carry the same pixels, geometry, compression and predictor as
// the reference's; the strip layout (and therefore the bytes) is the writer's own.  Reads uncompressed and
// LZW-compressed (predictor 1 or 2) chunky 16-bit strip TIFFs of either byte order: our own products and the
// reference's.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace oiptiff {

struct Info {
    int64_t width = 0, height = 0, rows_per_strip = 0;
    int spp = 1, bits = 16, compression = 1, planar = 1, photometric = 1, predictor = 1;
    bool big = false, little = true;
    std::vector<uint64_t> strip_off, strip_cnt;
};

namespace detail {
struct Out {
    std::vector<uint8_t> b;
    void u16(uint16_t v) { b.push_back((uint8_t)v); b.push_back((uint8_t)(v >> 8)); }
    void u32(uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
    void u64(uint64_t v) { for (int i = 0; i < 8; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
};
struct Entry {
    uint16_t tag, type; // 3 SHORT, 4 LONG, 16 LONG8
    std::vector<uint64_t> v;
};
inline size_t type_size(int t) { return t == 3 ? 2 : (t == 4 ? 4 : (t == 16 ? 8 : 1)); }
} // namespace detail

// pixels: height x width x spp, u16, row-major, samples interleaved in FILE order (the caller applies cv::imwrite's
// channel swap).  photometric: 1 = BlackIsZero (1 sample), 2 = RGB (+ one unassociated-alpha extra sample when spp == 4)
enum { COMPRESS_NONE = 1, COMPRESS_LZW = 5 };
inline void write_u16_lzw(const std::string &path, const uint16_t *pixels, int64_t width, int64_t height, int spp, int photometric);

// Returns the file offset of the first pixel of an uncompressed file (0 for LZW).  pixels == nullptr (uncompressed
// only): header and IFD only -- the caller writes the height x width x spp samples itself from that offset on (row blocks
// written in parallel by a streaming pipeline).
inline uint64_t write_u16(const std::string &path, const uint16_t *pixels, int64_t width, int64_t height, int spp, int photometric,
                          int compression = COMPRESS_NONE)
{
    using namespace detail;
    if (width < 1 || height < 1 || (spp != 1 && spp != 4)) throw std::invalid_argument("tiff: unsupported geometry");
    if (compression == COMPRESS_LZW) {
        if (!pixels) throw std::invalid_argument("tiff: header-only mode needs an uncompressed file");
        write_u16_lzw(path, pixels, width, height, spp, photometric);
        return 0;
    }
    if (compression != COMPRESS_NONE) throw std::invalid_argument("tiff: unsupported compression");
    const uint64_t row_bytes = (uint64_t)width * spp * 2, data_bytes = row_bytes * (uint64_t)height;
    int64_t rps = (int64_t)((8u << 20) / row_bytes);
    if (rps < 1) rps = 1;
    if (rps > height) rps = height;
    const uint64_t n_strips = (uint64_t)((height + rps - 1) / rps);
    const bool big = data_bytes + n_strips * 16 + 4096 >= 0xFFFF0000ull;
    const int off_t = big ? 16 : 4;
    std::vector<Entry> e;
    e.push_back({256, 4, {(uint64_t)width}});
    e.push_back({257, 4, {(uint64_t)height}});
    e.push_back({258, 3, std::vector<uint64_t>((size_t)spp, 16)});
    e.push_back({259, 3, {1}});
    e.push_back({262, 3, {(uint64_t)photometric}});
    e.push_back({273, (uint16_t)off_t, std::vector<uint64_t>(n_strips, 0)}); // filled below
    e.push_back({277, 3, {(uint64_t)spp}});
    e.push_back({278, 4, {(uint64_t)rps}});
    e.push_back({279, (uint16_t)off_t, std::vector<uint64_t>(n_strips, 0)});
    e.push_back({284, 3, {1}});
    if (spp == 4) e.push_back({338, 3, {2}});
    e.push_back({339, 3, std::vector<uint64_t>((size_t)spp, 1)});
    // layout: header | IFD | out-of-line values | pad | pixel data
    const size_t hdr = big ? 16 : 8, ent = big ? 20 : 12, inl = big ? 8 : 4;
    const size_t ifd = hdr, ifd_bytes = (big ? 8 : 2) + e.size() * ent + (big ? 8 : 4);
    size_t extra = ifd + ifd_bytes;
    std::vector<size_t> where(e.size(), 0);
    for (size_t i = 0; i < e.size(); ++i) {
        const size_t nb = e[i].v.size() * type_size(e[i].type);
        if (nb > inl) { where[i] = extra; extra += (nb + 7) & ~(size_t)7; }
    }
    const uint64_t data0 = (extra + 255) & ~(uint64_t)255;
    for (uint64_t s = 0; s < n_strips; ++s) {
        const uint64_t r0 = s * (uint64_t)rps, nr = std::min<uint64_t>((uint64_t)rps, (uint64_t)height - r0);
        e[5].v[s] = data0 + r0 * row_bytes;
        e[8].v[s] = nr * row_bytes;
    }
    Out o;
    o.b.reserve(data0);
    o.b.push_back('I'); o.b.push_back('I');
    if (big) { o.u16(43); o.u16(8); o.u16(0); o.u64(ifd); } else { o.u16(42); o.u32((uint32_t)ifd); }
    if (big) o.u64(e.size()); else o.u16((uint16_t)e.size());
    auto put = [&](Out &dst, int type, uint64_t v) { if (type == 3) dst.u16((uint16_t)v); else if (type == 4) dst.u32((uint32_t)v); else dst.u64(v); };
    for (size_t i = 0; i < e.size(); ++i) {
        o.u16(e[i].tag); o.u16(e[i].type);
        if (big) o.u64(e[i].v.size()); else o.u32((uint32_t)e[i].v.size());
        const size_t at = o.b.size();
        if (where[i]) { if (big) o.u64(where[i]); else o.u32((uint32_t)where[i]); }
        else {
            for (uint64_t v : e[i].v) put(o, e[i].type, v);
            while (o.b.size() < at + inl) o.b.push_back(0);
        }
    }
    if (big) o.u64(0); else o.u32(0); // no next IFD
    for (size_t i = 0; i < e.size(); ++i) {
        if (!where[i]) continue;
        while (o.b.size() < where[i]) o.b.push_back(0);
        for (uint64_t v : e[i].v) put(o, e[i].type, v);
    }
    while (o.b.size() < data0) o.b.push_back(0);
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("tiff: cannot create " + path);
    bool ok = fwrite(o.b.data(), 1, o.b.size(), f) == o.b.size();
    const uint8_t *p = reinterpret_cast<const uint8_t *>(pixels);
    for (uint64_t done = 0; ok && p && done < data_bytes;) {
        const size_t n = (size_t)std::min<uint64_t>(data_bytes - done, 64u << 20);
        ok = fwrite(p + done, 1, n, f) == n;
        done += n;
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) throw std::runtime_error("tiff: write failed: " + path);
    return data0;
}

namespace detail {
// TIFF LZW encoder, the counterpart of lzw_decode below and of libtiff's LZWEncode / LZWPostEncode: ClearCode first,
// codes MSB-first, the width grows when the next free code exceeds 2^width - 1 (which a decoder, one entry behind,
// sees "one early"), the table is cleared at 4094 entries, EndOfInformation last.
inline void lzw_encode(const uint8_t *src, size_t n, std::vector<uint8_t> &out)
{
    constexpr int HSIZE = 9001, HSHIFT = 13 - 8;
    struct Slot { int32_t key; uint16_t code; };
    std::vector<Slot> tab(HSIZE);
    auto reset = [&]() { for (auto &t : tab) t.key = -1; };
    uint64_t acc = 0;
    int nacc = 0, width = 9, next = 258;
    auto put = [&](int code) {
        acc = (acc << width) | (uint32_t)code;
        nacc += width;
        while (nacc >= 8) { out.push_back((uint8_t)(acc >> (nacc - 8))); nacc -= 8; }
    };
    reset();
    put(256);
    if (n == 0) { put(257); if (nacc) out.push_back((uint8_t)(acc << (8 - nacc))); return; }
    int w = src[0];
    for (size_t i = 1; i < n; ++i) {
        const int c = src[i];
        const int32_t key = (c << 12) + w;
        int h = (c << HSHIFT) ^ w;
        bool found = false;
        if (tab[h].key == key) { w = tab[h].code; continue; }
        if (tab[h].key >= 0) { // secondary probe, like libtiff
            int disp = h == 0 ? 1 : HSIZE - h;
            do {
                if ((h -= disp) < 0) h += HSIZE;
                if (tab[h].key == key) { w = tab[h].code; found = true; break; }
            } while (tab[h].key >= 0);
        }
        if (found) continue;
        put(w);
        tab[h].key = key;
        tab[h].code = (uint16_t)next++;
        if (next == 4094) { put(256); reset(); next = 258; width = 9; }
        else if (next > (1 << width) - 1) ++width;
        w = c;
    }
    put(w);
    ++next;
    if (next == 4094) { put(256); width = 9; }
    else if (next > (1 << width) - 1) ++width;
    put(257);
    if (nacc) out.push_back((uint8_t)(acc << (8 - nacc)));
}
} // namespace detail

// COMPRESS=LZW, PREDICTOR=2 (horizontal differencing of the 16-bit samples, per channel), strips of ~1 MiB compressed on
// all host cores and written in order; the IFD follows the data (its position is patched into the header)
inline void write_u16_lzw(const std::string &path, const uint16_t *pixels, int64_t width, int64_t height, int spp, int photometric)
{
    using namespace detail;
    const uint64_t row_bytes = (uint64_t)width * spp * 2, data_bytes = row_bytes * (uint64_t)height;
    int64_t rps = (int64_t)((1u << 20) / row_bytes);
    if (rps < 1) rps = 1;
    if (rps > height) rps = height;
    const uint64_t n_strips = (uint64_t)((height + rps - 1) / rps);
    const bool big = data_bytes + data_bytes / 2 + n_strips * 16 + 4096 >= 0xFFFF0000ull; // LZW can expand noise by ~40 %
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("tiff: cannot create " + path);
    bool ok = true;
    {
        Out h;
        h.b.push_back('I'); h.b.push_back('I');
        if (big) { h.u16(43); h.u16(8); h.u16(0); h.u64(0); } else { h.u16(42); h.u32(0); }
        ok = fwrite(h.b.data(), 1, h.b.size(), f) == h.b.size();
    }
    uint64_t pos = big ? 16 : 8;
    std::vector<uint64_t> offs(n_strips), cnts(n_strips);
    unsigned n_thr = std::thread::hardware_concurrency();
    if (n_thr < 1) n_thr = 1;
    if (n_thr > 32) n_thr = 32;
    const uint64_t batch = (uint64_t)n_thr * 2;
    std::vector<std::vector<uint8_t>> bufs(batch);
    for (uint64_t s0 = 0; ok && s0 < n_strips; s0 += batch) {
        const uint64_t nb = std::min<uint64_t>(batch, n_strips - s0);
        auto work = [&](uint64_t k) {
            const uint64_t s = s0 + k, r0 = s * (uint64_t)rps, nr = std::min<uint64_t>((uint64_t)rps, (uint64_t)height - r0);
            std::vector<uint16_t> diff((size_t)(nr * row_bytes / 2));
            const uint64_t rs = (uint64_t)width * spp;
            for (uint64_t y = 0; y < nr; ++y) {
                const uint16_t *row = pixels + (r0 + y) * rs;
                uint16_t *d = diff.data() + y * rs;
                for (uint64_t i = 0; i < rs; ++i) d[i] = i < (uint64_t)spp ? row[i] : (uint16_t)(row[i] - row[i - spp]);
            }
            bufs[k].clear();
            bufs[k].reserve((size_t)(nr * row_bytes / 2));
            lzw_encode(reinterpret_cast<const uint8_t *>(diff.data()), (size_t)(nr * row_bytes), bufs[k]);
        };
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < n_thr; ++t)
            pool.emplace_back([&, t]() { for (uint64_t k = t; k < nb; k += n_thr) work(k); });
        for (auto &th : pool) th.join();
        for (uint64_t k = 0; ok && k < nb; ++k) {
            offs[s0 + k] = pos;
            cnts[s0 + k] = bufs[k].size();
            ok = fwrite(bufs[k].data(), 1, bufs[k].size(), f) == bufs[k].size();
            pos += bufs[k].size();
            if (ok && (pos & 1)) { ok = fputc(0, f) != EOF; ++pos; } // strips start on even offsets
        }
    }
    // ---- IFD after the data
    const int off_t = big ? 16 : 4;
    std::vector<Entry> e;
    e.push_back({256, 4, {(uint64_t)width}});
    e.push_back({257, 4, {(uint64_t)height}});
    e.push_back({258, 3, std::vector<uint64_t>((size_t)spp, 16)});
    e.push_back({259, 3, {(uint64_t)COMPRESS_LZW}});
    e.push_back({262, 3, {(uint64_t)photometric}});
    e.push_back({273, (uint16_t)off_t, offs});
    e.push_back({277, 3, {(uint64_t)spp}});
    e.push_back({278, 4, {(uint64_t)rps}});
    e.push_back({279, (uint16_t)off_t, cnts});
    e.push_back({284, 3, {1}});
    e.push_back({317, 3, {2}});
    if (spp == 4) e.push_back({338, 3, {2}});
    e.push_back({339, 3, std::vector<uint64_t>((size_t)spp, 1)});
    const size_t ent = big ? 20 : 12, inl = big ? 8 : 4;
    const uint64_t ifd = (pos + 7) & ~(uint64_t)7;
    const size_t ifd_bytes = (big ? 8 : 2) + e.size() * ent + (big ? 8 : 4);
    uint64_t extra = ifd + ifd_bytes;
    std::vector<uint64_t> where(e.size(), 0);
    for (size_t i = 0; i < e.size(); ++i) {
        const size_t nb = e[i].v.size() * type_size(e[i].type);
        if (nb > inl) { where[i] = extra; extra += (nb + 7) & ~(size_t)7; }
    }
    if (!big && extra >= 0xFFFFFFF0ull) { fclose(f); throw std::runtime_error("tiff: classic TIFF overflow: " + path); }
    Out o;
    for (uint64_t q = pos; q < ifd; ++q) o.b.push_back(0);
    if (big) o.u64(e.size()); else o.u16((uint16_t)e.size());
    auto put = [&](Out &dst, int type, uint64_t v) { if (type == 3) dst.u16((uint16_t)v); else if (type == 4) dst.u32((uint32_t)v); else dst.u64(v); };
    for (size_t i = 0; i < e.size(); ++i) {
        o.u16(e[i].tag); o.u16(e[i].type);
        if (big) o.u64(e[i].v.size()); else o.u32((uint32_t)e[i].v.size());
        const size_t at = o.b.size();
        if (where[i]) { if (big) o.u64(where[i]); else o.u32((uint32_t)where[i]); }
        else {
            for (uint64_t v : e[i].v) put(o, e[i].type, v);
            while (o.b.size() < at + inl) o.b.push_back(0);
        }
    }
    if (big) o.u64(0); else o.u32(0);
    for (size_t i = 0; i < e.size(); ++i) {
        if (!where[i]) continue;
        while (pos + o.b.size() < where[i]) o.b.push_back(0);
        for (uint64_t v : e[i].v) put(o, e[i].type, v);
    }
    ok = ok && fwrite(o.b.data(), 1, o.b.size(), f) == o.b.size();
    Out hp;
    if (big) hp.u64(ifd); else hp.u32((uint32_t)ifd);
    ok = ok && fseeko(f, big ? 8 : 4, SEEK_SET) == 0 && fwrite(hp.b.data(), 1, hp.b.size(), f) == hp.b.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) throw std::runtime_error("tiff: write failed: " + path);
}

namespace detail {
struct In {
    FILE *f;
    bool little;
    uint64_t rd(int n)
    {
        uint8_t b[8];
        if (fread(b, 1, (size_t)n, f) != (size_t)n) throw std::runtime_error("tiff: truncated file");
        uint64_t v = 0;
        for (int i = 0; i < n; ++i) v |= (uint64_t)b[little ? i : n - 1 - i] << (8 * i);
        return v;
    }
};
} // namespace detail

inline Info read_info(const std::string &path)
{
    using namespace detail;
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("tiff: cannot open " + path);
    Info I;
    try {
        char bo[2];
        if (fread(bo, 1, 2, f) != 2 || (bo[0] != 'I' && bo[0] != 'M') || bo[0] != bo[1]) throw std::runtime_error("tiff: not a TIFF file: " + path);
        In in{f, bo[0] == 'I'};
        I.little = in.little;
        const uint64_t magic = in.rd(2);
        uint64_t ifd;
        if (magic == 42) ifd = in.rd(4);
        else if (magic == 43) { I.big = true; in.rd(2); in.rd(2); ifd = in.rd(8); }
        else throw std::runtime_error("tiff: not a TIFF file: " + path);
        fseeko(f, (off_t)ifd, SEEK_SET);
        const uint64_t n = I.big ? in.rd(8) : in.rd(2);
        for (uint64_t k = 0; k < n; ++k) {
            fseeko(f, (off_t)(ifd + (I.big ? 8 : 2) + k * (I.big ? 20 : 12)), SEEK_SET);
            const int tag = (int)in.rd(2), type = (int)in.rd(2);
            const uint64_t cnt = I.big ? in.rd(8) : in.rd(4);
            const size_t ts = type == 3 ? 2 : (type == 4 ? 4 : (type == 16 ? 8 : (type == 1 ? 1 : 0)));
            auto values = [&]() {
                std::vector<uint64_t> v;
                if (!ts) return v;
                if (cnt * ts > (I.big ? 8u : 4u)) fseeko(f, (off_t)(I.big ? in.rd(8) : in.rd(4)), SEEK_SET);
                for (uint64_t i = 0; i < cnt; ++i) v.push_back(in.rd((int)ts));
                return v;
            };
            switch (tag) {
            case 256: I.width = (int64_t)values().at(0); break;
            case 257: I.height = (int64_t)values().at(0); break;
            case 258: I.bits = (int)values().at(0); break;
            case 259: I.compression = (int)values().at(0); break;
            case 262: I.photometric = (int)values().at(0); break;
            case 273: I.strip_off = values(); break;
            case 277: I.spp = (int)values().at(0); break;
            case 278: I.rows_per_strip = (int64_t)values().at(0); break;
            case 279: I.strip_cnt = values(); break;
            case 284: I.planar = (int)values().at(0); break;
            case 317: I.predictor = (int)values().at(0); break;
            case 322: case 323: case 324: case 325: throw std::runtime_error("tiff: tiled TIFF is not supported: " + path);
            default: break;
            }
        }
    } catch (...) {
        fclose(f);
        throw;
    }
    fclose(f);
    if (I.rows_per_strip <= 0 || I.rows_per_strip > I.height) I.rows_per_strip = I.height;
    return I;
}

namespace detail {
// TIFF LZW (compression 5): MSB-first codes of 9..12 bits, 256 = clear, 257 = end of information, the code width grows
// one entry early (libtiff's "early change"), every strip is its own stream.  Returns the bytes produced.
inline size_t lzw_decode(const uint8_t *src, size_t n_src, uint8_t *dst, size_t n_dst)
{
    struct Entry { uint16_t prefix; uint8_t last, first; uint32_t len; };
    std::vector<Entry> tab(4096);
    for (int i = 0; i < 256; ++i) tab[i] = {0xFFFF, (uint8_t)i, (uint8_t)i, 1};
    int next = 258, width = 9;
    int prev = -1;
    uint64_t acc = 0;
    int nbits = 0;
    size_t ip = 0, op = 0;
    for (;;) {
        while (nbits < width && ip < n_src) { acc = (acc << 8) | src[ip++]; nbits += 8; }
        if (nbits < width) break;
        const int code = (int)((acc >> (nbits - width)) & ((1u << width) - 1));
        nbits -= width;
        if (code == 257) break;
        if (code == 256) { next = 258; width = 9; prev = -1; continue; }
        int cur = code;
        if (prev < 0) {
            if (code >= 256) throw std::runtime_error("tiff: corrupt LZW stream");
        } else {
            if (code > next) throw std::runtime_error("tiff: corrupt LZW stream");
            if (next < 4096) { // new entry = string(prev) + first byte of string(code) (or of string(prev) when code == next)
                const uint8_t fb = code < next ? tab[code].first : tab[prev].first;
                tab[next] = {(uint16_t)prev, fb, tab[prev].first, tab[prev].len + 1};
                ++next;
            }
        }
        const uint32_t len = tab[cur].len;
        if (op + len > n_dst) { // the last strip may carry padding: stop at the raster's end
            std::vector<uint8_t> tmp(len);
            int c = cur;
            for (uint32_t k = len; k-- > 0;) { tmp[k] = tab[c].last; c = tab[c].prefix; }
            const size_t take = n_dst - op;
            memcpy(dst + op, tmp.data(), take);
            op += take;
            break;
        }
        int c = cur;
        for (uint32_t k = len; k-- > 0;) { dst[op + k] = tab[c].last; c = tab[c].prefix; }
        op += len;
        prev = cur;
        if (next + 1 >= (1 << width) && width < 12) ++width; // early change
    }
    return op;
}
} // namespace detail

// the whole raster, u16 host order, samples interleaved in file order.  Uncompressed or LZW (what cv::imwrite and the
// reference's GTiff options produce, ref imageop.h:470-474), predictor 1 or 2, chunky, 16-bit, strips.
inline void read_u16(const std::string &path, const Info &I, uint16_t *dst)
{
    if (I.compression != 1 && I.compression != 5)
        throw std::runtime_error("tiff: compressed TIFF (compression " + std::to_string(I.compression) +
                                 ") is not supported by this build -- re-save it uncompressed or with LZW: " + path);
    if (I.bits != 16 || I.planar != 1 || I.strip_off.empty() || I.strip_off.size() != I.strip_cnt.size() ||
        (I.predictor != 1 && I.predictor != 2))
        throw std::runtime_error("tiff: only 16-bit chunky strip TIFF (predictor 1 or 2) is supported: " + path);
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("tiff: cannot open " + path);
    const uint64_t row_bytes = (uint64_t)I.width * I.spp * 2;
    bool ok = true;
    std::vector<uint8_t> comp;
    try {
        for (size_t s = 0; ok && s < I.strip_off.size(); ++s) {
            const uint64_t r0 = (uint64_t)s * (uint64_t)I.rows_per_strip;
            if (r0 >= (uint64_t)I.height) break;
            const uint64_t want = std::min<uint64_t>((uint64_t)I.rows_per_strip, (uint64_t)I.height - r0) * row_bytes;
            uint8_t *out = reinterpret_cast<uint8_t *>(dst) + r0 * row_bytes;
            if (I.compression == 1) {
                ok = I.strip_cnt[s] >= want && fseeko(f, (off_t)I.strip_off[s], SEEK_SET) == 0 && fread(out, 1, (size_t)want, f) == (size_t)want;
            } else {
                comp.resize((size_t)I.strip_cnt[s]);
                ok = fseeko(f, (off_t)I.strip_off[s], SEEK_SET) == 0 && fread(comp.data(), 1, comp.size(), f) == comp.size() &&
                     detail::lzw_decode(comp.data(), comp.size(), out, (size_t)want) == (size_t)want;
            }
        }
    } catch (...) {
        fclose(f);
        throw;
    }
    fclose(f);
    if (!ok) throw std::runtime_error("tiff: truncated or inconsistent strips: " + path);
    const uint64_t n = (uint64_t)I.width * I.height * I.spp;
    if (!I.little)
        for (uint64_t i = 0; i < n; ++i) dst[i] = (uint16_t)((dst[i] >> 8) | (dst[i] << 8));
    if (I.predictor == 2) { // horizontal differencing per sample, row by row (on the 16-bit values)
        const uint64_t rs = (uint64_t)I.width * I.spp;
        for (int64_t y = 0; y < I.height; ++y) {
            uint16_t *row = dst + (uint64_t)y * rs;
            for (uint64_t i = (uint64_t)I.spp; i < rs; ++i) row[i] = (uint16_t)(row[i] + row[i - I.spp]);
        }
    }
}

} // namespace oiptiff
