// Image side of the asset importer (include/kjb_asset.h): encoded bytes -> RGBA8 -> mip chain.
//
// Mirrors kajiya-asset/src/image.rs: `LoadImage::run` (:62-98: image::load_from_memory + to_rgba8) and the uncompressed branch of
// `CreateGpuImage::process_rgba8` (:130-283).  The decoders the reference reaches through the `image` crate (0.23.14: png 0.16.8,
// jpeg-decoder 0.1.22) are not part of /root/reference; what is restated here is the format (PNG: RFC 2083 + zlib/deflate RFC 1950/1951,
// bit-exact by construction; JPEG: ITU T.81 baseline/progressive Huffman with the integer IDCT, "fancy" chroma upsampling and float
// YCbCr conversion jpeg-decoder inherited from stb_image) and `imageops::resize(.., Lanczos3)` (vertical pass into f32, horizontal pass,
// weights normalised per output texel, round-to-nearest into u8).
#include "../../../include/kjb_asset.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace kjb_asset_detail {
void set_error(const std::string& m);

// ------------------------------------------------------------------------------------------------ inflate (RFC 1951) + zlib (RFC 1950)
namespace {
struct BitReader {
    const uint8_t* p; size_t n, pos = 0; uint64_t acc = 0; int cnt = 0; bool overrun = false;
    BitReader(const uint8_t* p_, size_t n_) : p(p_), n(n_) {}
    void fill() { while (cnt <= 56) { if (pos < n) acc |= uint64_t(p[pos++]) << cnt; else if (pos++ > n + 8) overrun = true; cnt += 8; } }
    uint32_t bits(int k) { if (k == 0) return 0; if (cnt < k) fill(); const uint32_t v = uint32_t(acc & ((1ull << k) - 1)); acc >>= k; cnt -= k; return v; }
    void align() { const int r = cnt & 7; acc >>= r; cnt -= r; }
};
struct Huff {
    uint16_t count[16] = {0}; uint16_t symbol[320];
    bool build(const uint8_t* lengths, int n) {
        memset(count, 0, sizeof count);
        for (int i = 0; i < n; ++i) count[lengths[i]]++;
        count[0] = 0;
        int left = 1;
        for (int l = 1; l < 16; ++l) { left <<= 1; left -= count[l]; if (left < 0) return false; }
        uint16_t offs[16]; offs[1] = 0;
        for (int l = 1; l < 15; ++l) offs[l + 1] = offs[l] + count[l];
        for (int i = 0; i < n; ++i) if (lengths[i]) symbol[offs[lengths[i]]++] = uint16_t(i);
        return true;
    }
    int decode(BitReader& br) const {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l < 16; ++l) {
            code |= int(br.bits(1));
            const int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c; first += c; first <<= 1; code <<= 1;
        }
        return -1;
    }
};
const uint16_t LEN_BASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const uint8_t LEN_EXTRA[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const uint16_t DIST_BASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
const uint8_t DIST_EXTRA[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

bool inflate_raw(BitReader& br, std::vector<uint8_t>& out, size_t size_hint) {
    out.reserve(size_hint);
    for (;;) {
        const uint32_t last = br.bits(1), type = br.bits(2);
        if (type == 0) {
            br.align();
            const uint32_t len = br.bits(16), nlen = br.bits(16);
            if ((len ^ 0xffffu) != nlen) return false;
            for (uint32_t i = 0; i < len; ++i) out.push_back(uint8_t(br.bits(8)));
        } else if (type == 1 || type == 2) {
            Huff lit, dist;
            if (type == 1) {
                uint8_t l[320];
                for (int i = 0; i < 144; ++i) l[i] = 8;
                for (int i = 144; i < 256; ++i) l[i] = 9;
                for (int i = 256; i < 280; ++i) l[i] = 7;
                for (int i = 280; i < 288; ++i) l[i] = 8;
                lit.build(l, 288);
                for (int i = 0; i < 30; ++i) l[i] = 5;
                dist.build(l, 30);
            } else {
                const int nlen = int(br.bits(5)) + 257, ndist = int(br.bits(5)) + 1, ncode = int(br.bits(4)) + 4;
                if (nlen > 286 || ndist > 30) return false;
                static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                uint8_t l[320] = {0};
                for (int i = 0; i < ncode; ++i) l[order[i]] = uint8_t(br.bits(3));
                Huff cl; if (!cl.build(l, 19)) return false;
                int idx = 0; uint8_t ll[320] = {0};
                while (idx < nlen + ndist) {
                    const int sym = cl.decode(br);
                    if (sym < 0) return false;
                    if (sym < 16) { ll[idx++] = uint8_t(sym); continue; }
                    int rep, val = 0;
                    if (sym == 16) { if (idx == 0) return false; val = ll[idx - 1]; rep = 3 + int(br.bits(2)); }
                    else if (sym == 17) rep = 3 + int(br.bits(3));
                    else rep = 11 + int(br.bits(7));
                    if (idx + rep > nlen + ndist) return false;
                    while (rep--) ll[idx++] = uint8_t(val);
                }
                if (ll[256] == 0) return false;
                if (!lit.build(ll, nlen)) return false;
                dist.build(ll + nlen, ndist);   // incomplete distance codes are legal (single-code case)
            }
            for (;;) {
                const int sym = lit.decode(br);
                if (sym < 0 || br.overrun) return false;
                if (sym < 256) { out.push_back(uint8_t(sym)); continue; }
                if (sym == 256) break;
                const int li = sym - 257; if (li >= 29) return false;
                const uint32_t len = LEN_BASE[li] + br.bits(LEN_EXTRA[li]);
                const int ds = dist.decode(br); if (ds < 0 || ds >= 30) return false;
                const size_t d = DIST_BASE[ds] + br.bits(DIST_EXTRA[ds]);
                if (d > out.size()) return false;
                const size_t start = out.size() - d;
                for (uint32_t i = 0; i < len; ++i) out.push_back(out[start + i]);
            }
        } else return false;
        if (br.overrun) return false;
        if (last) return true;
    }
}
bool zlib_decompress(const uint8_t* p, size_t n, std::vector<uint8_t>& out, size_t size_hint) {
    if (n < 6) return false;
    const uint32_t cmf = p[0], flg = p[1];
    if ((cmf & 15) != 8 || ((cmf << 8) | flg) % 31 != 0 || (flg & 0x20)) return false;
    BitReader br(p + 2, n - 2);
    if (!inflate_raw(br, out, size_hint)) return false;
    uint32_t a = 1, b = 0;
    for (size_t i = 0; i < out.size();) { const size_t e = std::min(out.size(), i + 5552); for (; i < e; ++i) { a += out[i]; b += a; } a %= 65521; b %= 65521; }
    br.align();
    const size_t tail = 2 + br.pos - size_t(br.cnt / 8);
    if (tail + 4 > n) return false;
    const uint32_t want = (uint32_t(p[tail]) << 24) | (uint32_t(p[tail + 1]) << 16) | (uint32_t(p[tail + 2]) << 8) | p[tail + 3];
    return want == ((b << 16) | a);
}

// ------------------------------------------------------------------------------------------------ PNG
uint32_t be32(const uint8_t* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
uint32_t crc32_of(const uint8_t* p, size_t n) {
    static uint32_t table[256]; static bool init = false;
    if (!init) { for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; } init = true; }
    uint32_t c = 0xffffffffu;
    for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xff] ^ (c >> 8);
    return c ^ 0xffffffffu;
}
int paeth(int a, int b, int c) { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); }
bool unfilter(uint8_t* data, size_t rows, size_t stride, size_t bpp) {   // data: rows * (1 + stride)
    std::vector<uint8_t> zero(stride, 0);
    const uint8_t* prev = zero.data();
    for (size_t y = 0; y < rows; ++y) {
        uint8_t* row = data + y * (stride + 1);
        const uint8_t ft = row[0]; uint8_t* cur = row + 1;
        switch (ft) {
            case 0: break;
            case 1: for (size_t i = bpp; i < stride; ++i) cur[i] = uint8_t(cur[i] + cur[i - bpp]); break;
            case 2: for (size_t i = 0; i < stride; ++i) cur[i] = uint8_t(cur[i] + prev[i]); break;
            case 3: for (size_t i = 0; i < stride; ++i) cur[i] = uint8_t(cur[i] + (((i >= bpp ? cur[i - bpp] : 0) + prev[i]) >> 1)); break;
            case 4: for (size_t i = 0; i < stride; ++i) cur[i] = uint8_t(cur[i] + paeth(i >= bpp ? cur[i - bpp] : 0, prev[i], i >= bpp ? prev[i - bpp] : 0)); break;
            default: return false;
        }
        prev = cur;
    }
    return true;
}
uint8_t narrow16(uint32_t v16) { return uint8_t((v16 + 128u) / 257u); }   // image 0.23 u16 -> u8 channel conversion (to_rgba8 on 16-bit buffers)

bool decode_png(const uint8_t* p, size_t n, std::vector<uint8_t>& rgba, uint32_t& W, uint32_t& H) {
    size_t off = 8;
    uint32_t depth = 0, ctype = 0, interlace = 0; bool have_ihdr = false;
    std::vector<uint8_t> idat, plte, trns;
    while (off + 12 <= n) {
        const uint32_t len = be32(p + off); const uint8_t* type = p + off + 4;
        if (off + 12 + size_t(len) > n) { set_error("png: truncated chunk"); return false; }
        if (crc32_of(type, 4 + size_t(len)) != be32(p + off + 8 + len)) { set_error("png: chunk CRC mismatch"); return false; }
        const uint8_t* d = p + off + 8;
        if (!memcmp(type, "IHDR", 4) && len == 13) {
            W = be32(d); H = be32(d + 4); depth = d[8]; ctype = d[9]; interlace = d[12]; have_ihdr = true;
            if (d[10] != 0 || d[11] != 0 || interlace > 1) { set_error("png: unsupported compression/filter/interlace method"); return false; }
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(d, d + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(d, d + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), d, d + len);
        else if (!memcmp(type, "IEND", 4)) break;
        off += 12 + size_t(len);
    }
    if (!have_ihdr || W == 0 || H == 0 || W > (1u << 16) || H > (1u << 16)) { set_error("png: missing or bad IHDR"); return false; }
    const uint32_t channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    const bool depth_ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) || (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8))
                          || ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
    if (!channels || !depth_ok) { set_error("png: bad colour type / bit depth"); return false; }
    if (ctype == 3 && plte.size() < 3) { set_error("png: palette image without PLTE"); return false; }
    const uint32_t bits_pp = channels * depth;
    const size_t bpp = std::max<size_t>(1, bits_pp / 8);
    auto row_bytes = [&](uint32_t w) { return (size_t(w) * bits_pp + 7) / 8; };

    // pass geometry: one pass for non-interlaced, seven for Adam7
    struct Pass { uint32_t x0, y0, dx, dy; };
    static const Pass ADAM7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const Pass WHOLE = {0, 0, 1, 1};
    const int npass = interlace ? 7 : 1;
    size_t raw_size = 0;
    for (int k = 0; k < npass; ++k) {
        const Pass& ps = interlace ? ADAM7[k] : WHOLE;
        const uint32_t pw = (W > ps.x0) ? (W - ps.x0 + ps.dx - 1) / ps.dx : 0, ph = (H > ps.y0) ? (H - ps.y0 + ps.dy - 1) / ps.dy : 0;
        if (pw && ph) raw_size += size_t(ph) * (1 + row_bytes(pw));
    }
    // deflate cannot expand by more than 1032:1, so a header that promises more than the IDAT stream could hold is corrupt (and must not drive the allocations below)
    if (raw_size > idat.size() * 1032 + 1024) { set_error("png: image data stream too short for the header's dimensions"); return false; }
    std::vector<uint8_t> raw;
    if (!zlib_decompress(idat.data(), idat.size(), raw, raw_size) || raw.size() < raw_size) { set_error("png: corrupt image data stream"); return false; }

    rgba.assign(size_t(W) * H * 4, 255);
    // sample fetch from an unfiltered row
    auto sample = [&](const uint8_t* row, uint32_t x, uint32_t c) -> uint32_t {
        if (depth == 8) return row[size_t(x) * channels + c];
        if (depth == 16) { const uint8_t* q = row + (size_t(x) * channels + c) * 2; return (uint32_t(q[0]) << 8) | q[1]; }
        const uint32_t bit = x * depth; return (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1);
    };
    const uint32_t gray_key = (ctype == 0 && trns.size() >= 2) ? ((uint32_t(trns[0]) << 8) | trns[1]) : 0xffffffffu;
    uint32_t rgb_key[3] = {0xffffffffu, 0, 0};
    if (ctype == 2 && trns.size() >= 6) for (int c = 0; c < 3; ++c) rgb_key[c] = (uint32_t(trns[2 * c]) << 8) | trns[2 * c + 1];
    size_t cursor = 0;
    for (int k = 0; k < npass; ++k) {
        const Pass& ps = interlace ? ADAM7[k] : WHOLE;
        const uint32_t pw = (W > ps.x0) ? (W - ps.x0 + ps.dx - 1) / ps.dx : 0, ph = (H > ps.y0) ? (H - ps.y0 + ps.dy - 1) / ps.dy : 0;
        if (!pw || !ph) continue;
        const size_t stride = row_bytes(pw);
        if (!unfilter(raw.data() + cursor, ph, stride, bpp)) { set_error("png: bad filter type"); return false; }
        for (uint32_t py = 0; py < ph; ++py) {
            const uint8_t* row = raw.data() + cursor + size_t(py) * (stride + 1) + 1;
            for (uint32_t px = 0; px < pw; ++px) {
                uint8_t* o = rgba.data() + (size_t(ps.y0 + py * ps.dy) * W + (ps.x0 + px * ps.dx)) * 4;
                if (ctype == 3) {
                    const uint32_t i = sample(row, px, 0);
                    if (size_t(i) * 3 + 2 < plte.size()) { o[0] = plte[i * 3]; o[1] = plte[i * 3 + 1]; o[2] = plte[i * 3 + 2]; } else { o[0] = o[1] = o[2] = 0; }
                    o[3] = i < trns.size() ? trns[i] : 255;
                } else if (ctype == 0 || ctype == 4) {
                    const uint32_t v = sample(row, px, 0);
                    // sub-byte grays are scaled to the full range (png EXPAND), 16-bit narrowed like to_rgba8 does
                    const uint8_t g = depth == 16 ? narrow16(v) : depth == 8 ? uint8_t(v) : uint8_t(v * (255u / ((1u << depth) - 1)));
                    o[0] = o[1] = o[2] = g;
                    if (ctype == 4) { const uint32_t a = sample(row, px, 1); o[3] = depth == 16 ? narrow16(a) : uint8_t(a); }
                    else o[3] = (v == gray_key) ? 0 : 255;
                } else {
                    uint32_t v[4] = {0, 0, 0, depth == 16 ? 65535u : 255u};
                    for (uint32_t c = 0; c < channels; ++c) v[c] = sample(row, px, c);
                    if (ctype == 2 && v[0] == rgb_key[0] && v[1] == rgb_key[1] && v[2] == rgb_key[2]) v[3] = 0;
                    for (int c = 0; c < 4; ++c) o[c] = depth == 16 ? narrow16(v[c]) : uint8_t(v[c]);
                }
            }
        }
        cursor += size_t(ph) * (stride + 1);
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ JPEG (ITU T.81, Huffman, 8-bit)
struct JHuff {
    uint8_t bits[17] = {0}; uint8_t vals[256] = {0};
    int32_t maxcode[18]; int32_t valptr[17]; uint16_t mincode[17]; bool present = false;
    void build() {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k; mincode[l] = uint16_t(code);
            code += bits[l]; k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff; present = true;
    }
};
struct JComp { int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0; int bw = 0, bh = 0; std::vector<int16_t> coef; int dc_pred = 0; };
struct JBits {
    const uint8_t* p; size_t n, pos; uint32_t acc = 0; int cnt = 0; bool hit_marker = false;
    JBits(const uint8_t* p_, size_t n_, size_t pos_) : p(p_), n(n_), pos(pos_) {}
    void fill() {
        while (cnt <= 24) {
            uint32_t b = 0;
            if (!hit_marker && pos < n) {
                b = p[pos];
                if (b == 0xff) {
                    const uint8_t nx = pos + 1 < n ? p[pos + 1] : 0xd9;
                    if (nx == 0) pos += 2; else { hit_marker = true; b = 0; }
                } else pos += 1;
            }
            acc |= b << (24 - cnt); cnt += 8;
        }
    }
    int bit() { if (cnt < 1) fill(); const int v = int(acc >> 31); acc <<= 1; cnt -= 1; return v; }
    int bits(int k) { if (k == 0) return 0; if (cnt < k) fill(); const int v = int(acc >> (32 - k)); acc <<= k; cnt -= k; return v; }
    void reset() { acc = 0; cnt = 0; hit_marker = false; }
};
int jdecode(JBits& br, const JHuff& h) {
    int code = 0;
    for (int l = 1; l <= 16; ++l) {
        code = (code << 1) | br.bit();
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= int(h.mincode[l])) return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    return -1;
}
int jextend(int v, int t) { return (t && v < (1 << (t - 1))) ? v - (1 << t) + 1 : v; }
const uint8_t ZIGZAG[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                            35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// integer IDCT of stb_image (which jpeg-decoder 0.1's idct.rs follows): 12-bit fixed point constants, rows then columns
inline int f2f(double x) { return int(x * 4096.0 + 0.5); }
inline uint8_t clamp_u8(int x) { return uint8_t(x < 0 ? 0 : x > 255 ? 255 : x); }
// dequantised coefficient, clamped to +-2^15: legitimate 8-bit JPEG coefficients stay below 2^12, so real files are untouched
inline int64_t jdeq(int16_t c, uint16_t q) { const int64_t v = int64_t(c) * int64_t(q); return v < -32768 ? -32768 : v > 32768 ? 32768 : v; }
inline uint8_t clamp_u8(int64_t x) { return uint8_t(x < 0 ? 0 : x > 255 ? 255 : x); }
void idct_block(const int16_t* in, const uint16_t* q, uint8_t* out, size_t out_stride) {
    int64_t val[64];
#define KJB_IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                                                             \
    int64_t t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;   /* 64-bit: hostile 16-bit DQT tables must not overflow (same values as int on real files) */ \
    p2 = s2; p3 = s6; p1 = (p2 + p3) * f2f(0.5411961); t2 = p1 + p3 * f2f(-1.847759065); t3 = p1 + p2 * f2f(0.765366865); \
    p2 = s0; p3 = s4; t0 = (p2 + p3) * 4096; t1 = (p2 - p3) * 4096;                                             \
    x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2;                                                     \
    t0 = s7; t1 = s5; t2 = s3; t3 = s1;                                                                         \
    p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2; p5 = (p3 + p4) * f2f(1.175875602);                  \
    t0 = t0 * f2f(0.298631336); t1 = t1 * f2f(2.053119869); t2 = t2 * f2f(3.072711026); t3 = t3 * f2f(1.501321110); \
    p1 = p5 + p1 * f2f(-0.899976223); p2 = p5 + p2 * f2f(-2.562915447); p3 = p3 * f2f(-1.961570560); p4 = p4 * f2f(-0.390180644); \
    t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;
    for (int i = 0; i < 8; ++i) {
        const int16_t* d = in + i; const uint16_t* qq = q + i; int64_t* v = val + i;
        if (d[8] == 0 && d[16] == 0 && d[24] == 0 && d[32] == 0 && d[40] == 0 && d[48] == 0 && d[56] == 0) {
            const int64_t dc = jdeq(d[0], qq[0]) * 4;
            v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dc;
        } else {
            KJB_IDCT_1D(jdeq(d[0], qq[0]), jdeq(d[8], qq[8]), jdeq(d[16], qq[16]), jdeq(d[24], qq[24]), jdeq(d[32], qq[32]), jdeq(d[40], qq[40]), jdeq(d[48], qq[48]), jdeq(d[56], qq[56]))
            x0 += 512; x1 += 512; x2 += 512; x3 += 512;
            v[0] = (x0 + t3) >> 10; v[56] = (x0 - t3) >> 10; v[8] = (x1 + t2) >> 10; v[48] = (x1 - t2) >> 10;
            v[16] = (x2 + t1) >> 10; v[40] = (x2 - t1) >> 10; v[24] = (x3 + t0) >> 10; v[32] = (x3 - t0) >> 10;
        }
    }
    for (int i = 0; i < 8; ++i) {
        const int64_t* v = val + i * 8; uint8_t* o = out + size_t(i) * out_stride;
        KJB_IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
        x0 += 65536 + (128 << 17); x1 += 65536 + (128 << 17); x2 += 65536 + (128 << 17); x3 += 65536 + (128 << 17);
        o[0] = clamp_u8((x0 + t3) >> 17); o[7] = clamp_u8((x0 - t3) >> 17); o[1] = clamp_u8((x1 + t2) >> 17); o[6] = clamp_u8((x1 - t2) >> 17);
        o[2] = clamp_u8((x2 + t1) >> 17); o[5] = clamp_u8((x2 - t1) >> 17); o[3] = clamp_u8((x3 + t0) >> 17); o[4] = clamp_u8((x3 - t0) >> 17);
    }
#undef KJB_IDCT_1D
}

struct JpegDecoder {
    const uint8_t* p; size_t n;
    uint16_t qt[4][64]; bool qt_present[4] = {false, false, false, false};
    JHuff dc[4], ac[4];
    std::vector<JComp> comps;
    int width = 0, height = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0, restart_interval = 0;
    bool progressive = false; int adobe_transform = -1; uint32_t eobrun = 0;

    bool fail(const char* m) { set_error(std::string("jpeg: ") + m); return false; }

    bool decode_block_baseline(JBits& br, JComp& c, int16_t* blk) {
        const int t = jdecode(br, dc[c.td]); if (t < 0 || t > 16) return fail("bad DC code");
        const int diff = t ? jextend(br.bits(t), t) : 0;
        c.dc_pred += diff; blk[0] = int16_t(c.dc_pred);
        for (int k = 1; k < 64;) {
            const int rs = jdecode(br, ac[c.ta]); if (rs < 0) return fail("bad AC code");
            const int r = rs >> 4, s = rs & 15;
            if (s == 0) { if (r == 15) { k += 16; continue; } break; }
            k += r; if (k > 63) return fail("AC run past block end");
            blk[ZIGZAG[k]] = int16_t(jextend(br.bits(s), s)); ++k;
        }
        return true;
    }
    bool decode_block_progressive(JBits& br, JComp& c, int16_t* blk, int ss, int se, int ah, int al) {
        if (ss == 0) {   // DC scan
            if (ah == 0) {
                const int t = jdecode(br, dc[c.td]); if (t < 0 || t > 16) return fail("bad DC code");
                const int diff = t ? jextend(br.bits(t), t) : 0;
                c.dc_pred += diff; blk[0] = int16_t(c.dc_pred * (1 << al));
            } else if (br.bit()) blk[0] = int16_t(blk[0] | (1 << al));
            return true;
        }
        if (ah == 0) {   // AC first pass
            if (eobrun) { --eobrun; return true; }
            for (int k = ss; k <= se;) {
                const int rs = jdecode(br, ac[c.ta]); if (rs < 0) return fail("bad AC code");
                const int r = rs >> 4, s = rs & 15;
                if (s == 0) {
                    if (r < 15) { eobrun = (1u << r) - 1; if (r) eobrun += uint32_t(br.bits(r)); break; }
                    k += 16; continue;
                }
                k += r; if (k > 63) return fail("AC run past block end");
                blk[ZIGZAG[k]] = int16_t(jextend(br.bits(s), s) * (1 << al)); ++k;
            }
            return true;
        }
        // AC refinement
        const int p1 = 1 << al, m1 = -1 * (1 << al);
        int k = ss;
        if (eobrun == 0) {
            for (; k <= se;) {
                const int rs = jdecode(br, ac[c.ta]); if (rs < 0) return fail("bad AC code");
                int r = rs >> 4; const int s = rs & 15; int value = 0;
                if (s == 0) {
                    if (r < 15) { eobrun = (1u << r); if (r) eobrun += uint32_t(br.bits(r)); break; }
                } else { if (s != 1) return fail("bad refinement size"); value = br.bit() ? p1 : m1; }
                while (k <= se) {
                    int16_t& co = blk[ZIGZAG[k++]];
                    if (co != 0) {
                        if (br.bit() && (co & p1) == 0) co = int16_t(co + (co > 0 ? p1 : m1));
                    } else {
                        if (r == 0) { if (value) co = int16_t(value); break; }
                        --r;
                    }
                }
            }
        }
        if (eobrun) {
            for (; k <= se; ++k) { int16_t& co = blk[ZIGZAG[k]]; if (co != 0 && br.bit() && (co & p1) == 0) co = int16_t(co + (co > 0 ? p1 : m1)); }
            --eobrun;
        }
        return true;
    }

    bool scan(size_t& pos, const std::vector<int>& order, int ss, int se, int ah, int al) {
        JBits br(p, n, pos);
        for (int ci : order) comps[ci].dc_pred = 0;
        eobrun = 0;
        const bool single = order.size() == 1;
        int todo = restart_interval ? restart_interval : 0x7fffffff;
        int total_units;
        if (single) {
            const JComp& c = comps[order[0]];
            const int cw = (width * c.h + hmax * 8 - 1) / (hmax * 8), ch = (height * c.v + vmax * 8 - 1) / (vmax * 8);   // non-interleaved: only blocks covering the image
            total_units = cw * ch;
            for (int u = 0; u < total_units; ++u) {
                const int bx = u % cw, by = u / cw;
                JComp& cc = comps[order[0]];
                int16_t* blk = cc.coef.data() + (size_t(by) * cc.bw + bx) * 64;
                if (!(progressive ? decode_block_progressive(br, cc, blk, ss, se, ah, al) : decode_block_baseline(br, cc, blk))) return false;
                if (--todo <= 0 && u + 1 < total_units) { if (!restart(br, order)) return false; todo = restart_interval; }
            }
        } else {
            total_units = mcux * mcuy;
            for (int u = 0; u < total_units; ++u) {
                const int mx = u % mcux, my = u / mcux;
                for (int ci : order) {
                    JComp& cc = comps[ci];
                    for (int y = 0; y < cc.v; ++y) for (int x = 0; x < cc.h; ++x) {
                        int16_t* blk = cc.coef.data() + (size_t(my * cc.v + y) * cc.bw + (mx * cc.h + x)) * 64;
                        if (!(progressive ? decode_block_progressive(br, cc, blk, ss, se, ah, al) : decode_block_baseline(br, cc, blk))) return false;
                    }
                }
                if (--todo <= 0 && u + 1 < total_units) { if (!restart(br, order)) return false; todo = restart_interval; }
            }
        }
        // leave `pos` at the marker that ended the entropy-coded segment
        size_t q = br.pos;
        if (!br.hit_marker) { while (q + 1 < n && !(p[q] == 0xff && p[q + 1] != 0 && !(p[q + 1] >= 0xd0 && p[q + 1] <= 0xd7))) ++q; }
        pos = q;
        return true;
    }
    bool restart(JBits& br, const std::vector<int>& order) {
        size_t q = br.pos;
        if (!br.hit_marker) { while (q + 1 < n && !(p[q] == 0xff && p[q + 1] >= 0xd0 && p[q + 1] <= 0xd7)) { if (p[q] == 0xff && p[q + 1] != 0 && p[q + 1] != 0xff) break; ++q; } }
        if (q + 1 >= n || p[q] != 0xff || p[q + 1] < 0xd0 || p[q + 1] > 0xd7) return fail("missing restart marker");
        br.pos = q + 2; br.reset();
        for (int ci : order) comps[ci].dc_pred = 0;
        eobrun = 0;
        return true;
    }

    bool run(std::vector<uint8_t>& rgba, uint32_t& W, uint32_t& H) {
        if (n < 4 || p[0] != 0xff || p[1] != 0xd8) return fail("missing SOI");
        size_t pos = 2; bool have_frame = false, done = false;
        while (!done && pos + 4 <= n) {
            if (p[pos] != 0xff) { ++pos; continue; }
            const uint8_t m = p[pos + 1];
            if (m == 0xff) { ++pos; continue; }
            if (m == 0xd9) break;
            if (m == 0x01 || (m >= 0xd0 && m <= 0xd7)) { pos += 2; continue; }
            const size_t len = (size_t(p[pos + 2]) << 8) | p[pos + 3];
            if (len < 2 || pos + 2 + len > n) return fail("truncated segment");
            const uint8_t* d = p + pos + 4; const size_t dl = len - 2;
            pos += 2 + len;
            if (m == 0xdb) {
                for (size_t o = 0; o < dl;) {
                    const int pq = d[o] >> 4, tq = d[o] & 15; ++o; if (tq > 3) return fail("bad DQT");
                    if (o + (pq ? 128 : 64) > dl) return fail("bad DQT");
                    for (int k = 0; k < 64; ++k) { qt[tq][ZIGZAG[k]] = pq ? uint16_t((d[o] << 8) | d[o + 1]) : d[o]; o += pq ? 2 : 1; }
                    qt_present[tq] = true;
                }
            } else if (m == 0xc4) {
                for (size_t o = 0; o + 17 <= dl;) {
                    const int tc = d[o] >> 4, th = d[o] & 15; if (th > 3 || tc > 1) return fail("bad DHT");
                    JHuff& h = tc ? ac[th] : dc[th];
                    int total = 0; for (int l = 1; l <= 16; ++l) { h.bits[l] = d[o + l]; total += h.bits[l]; }
                    o += 17; if (total > 256 || o + total > dl) return fail("bad DHT");
                    memcpy(h.vals, d + o, total); o += total; h.build();
                }
            } else if (m == 0xc0 || m == 0xc1 || m == 0xc2) {
                if (have_frame) return fail("multiple frames");
                progressive = m == 0xc2;
                if (dl < 6 || d[0] != 8) return fail("only 8-bit precision is supported");
                height = (d[1] << 8) | d[2]; width = (d[3] << 8) | d[4];
                const int nc = d[5];
                if (!width || !height) return fail("zero-sized frame");
                if (nc != 1 && nc != 3) return fail("only greyscale and 3-component images are supported");
                if (dl < size_t(6 + 3 * nc)) return fail("bad SOF");
                comps.resize(nc);
                for (int i = 0; i < nc; ++i) { comps[i].id = d[6 + 3 * i]; comps[i].h = d[7 + 3 * i] >> 4; comps[i].v = d[7 + 3 * i] & 15; comps[i].tq = d[8 + 3 * i];
                    if (comps[i].h < 1 || comps[i].h > 4 || comps[i].v < 1 || comps[i].v > 4 || comps[i].tq > 3) return fail("bad component spec"); }
                if (nc == 1) { comps[0].h = comps[0].v = 1; }
                hmax = vmax = 1; for (auto& c : comps) { hmax = std::max(hmax, c.h); vmax = std::max(vmax, c.v); }
                mcux = (width + 8 * hmax - 1) / (8 * hmax); mcuy = (height + 8 * vmax - 1) / (8 * vmax);
                {   // plausibility before sizing the coefficient planes: every 8x8 block costs at least one entropy-coded bit, so a frame
                    // cannot hold more blocks than 8x the bytes that are left (plus a hard cap); a tiny file with a 65535^2 SOF is refused here
                    size_t blocks = 0; for (auto& c : comps) blocks += size_t(mcux) * c.h * size_t(mcuy) * c.v;
                    if (blocks > (n - pos) * 8 + 64 || blocks > (size_t(1) << 24)) return fail("frame size is implausible for the amount of data");
                }
                for (auto& c : comps) { c.bw = mcux * c.h; c.bh = mcuy * c.v; c.coef.assign(size_t(c.bw) * c.bh * 64, 0); }
                have_frame = true;
            } else if (m >= 0xc3 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc) {
                return fail("unsupported coding process (lossless / arithmetic / hierarchical)");
            } else if (m == 0xdd) {
                if (dl < 2) return fail("bad DRI");
                restart_interval = (d[0] << 8) | d[1];
            } else if (m == 0xee) {
                if (dl >= 12 && !memcmp(d, "Adobe", 5)) adobe_transform = d[11];
            } else if (m == 0xda) {
                if (!have_frame) return fail("scan before frame header");
                if (dl < 1) return fail("bad SOS");
                const int ns = d[0]; if (ns < 1 || ns > int(comps.size()) || dl < size_t(4 + 2 * ns)) return fail("bad SOS");
                std::vector<int> order;
                for (int i = 0; i < ns; ++i) {
                    int ci = -1; for (size_t k = 0; k < comps.size(); ++k) if (comps[k].id == d[1 + 2 * i]) ci = int(k);
                    if (ci < 0) return fail("scan names an unknown component");
                    comps[ci].td = d[2 + 2 * i] >> 4; comps[ci].ta = d[2 + 2 * i] & 15;
                    if (comps[ci].td > 3 || comps[ci].ta > 3) return fail("bad table selector");
                    order.push_back(ci);
                }
                const int ss = d[1 + 2 * ns], se = d[2 + 2 * ns], ah = d[3 + 2 * ns] >> 4, al = d[3 + 2 * ns] & 15;
                if (progressive) { if (ss > se || se > 63 || (ss == 0 && se != 0) || (ss != 0 && ns != 1)) return fail("bad progressive scan parameters"); }
                for (int ci : order) {
                    if ((!progressive || ss == 0) && !(progressive && ah) && !dc[comps[ci].td].present) return fail("scan uses an undefined DC table");
                    if ((!progressive || ss != 0) && !ac[comps[ci].ta].present) return fail("scan uses an undefined AC table");
                }
                if (!scan(pos, order, progressive ? ss : 0, progressive ? se : 63, ah, al)) return false;
            }
        }
        if (!have_frame) return fail("no frame");
        for (auto& c : comps) if (!qt_present[c.tq]) return fail("component uses an undefined quantisation table");

        // reconstruct planes
        std::vector<std::vector<uint8_t>> planes(comps.size());
        for (size_t i = 0; i < comps.size(); ++i) {
            JComp& c = comps[i];
            const size_t stride = size_t(c.bw) * 8;
            planes[i].assign(stride * c.bh * 8, 0);
            for (int by = 0; by < c.bh; ++by) for (int bx = 0; bx < c.bw; ++bx)
                idct_block(c.coef.data() + (size_t(by) * c.bw + bx) * 64, qt[c.tq], planes[i].data() + size_t(by) * 8 * stride + size_t(bx) * 8, stride);
        }
        W = uint32_t(width); H = uint32_t(height);
        rgba.assign(size_t(W) * H * 4, 255);
        if (comps.size() == 1) {
            const size_t stride = size_t(comps[0].bw) * 8;
            for (uint32_t y = 0; y < H; ++y) for (uint32_t x = 0; x < W; ++x) { uint8_t* o = &rgba[(size_t(y) * W + x) * 4]; o[0] = o[1] = o[2] = planes[0][y * stride + x]; }
            return true;
        }
        // chroma upsampling per output row (jpeg-decoder upsampler.rs: H1V1 copy, H2V1 / H2V2 triangle filters after stb_image)
        std::vector<std::vector<uint8_t>> lines(3, std::vector<uint8_t>(size_t(W) + 2 * 8 * 4 + 16));
        for (uint32_t y = 0; y < H; ++y) {
            for (int i = 0; i < 3; ++i) {
                const JComp& c = comps[i];
                const size_t stride = size_t(c.bw) * 8;
                const int cw = (width * c.h + hmax - 1) / hmax, chh = (height * c.v + vmax - 1) / vmax;   // component extent in samples
                const int hs = hmax / c.h, vs = vmax / c.v;
                uint8_t* out = lines[i].data();
                if (hmax % c.h || vmax % c.v) return fail("fractional sampling ratios are not supported");
                if (hs == 1 && vs == 1) { memcpy(out, planes[i].data() + size_t(y) * stride, W); }
                else if (hs == 2 && vs == 1) {
                    const uint8_t* in = planes[i].data() + size_t(y) * stride;
                    if (cw == 1) { out[0] = out[1] = in[0]; }
                    else {
                        out[0] = in[0]; out[1] = uint8_t((in[0] * 3 + in[1] + 2) >> 2);
                        for (int x = 1; x < cw - 1; ++x) { const int s = 3 * in[x] + 2; out[2 * x] = uint8_t((s + in[x - 1]) >> 2); out[2 * x + 1] = uint8_t((s + in[x + 1]) >> 2); }
                        out[(cw - 1) * 2] = uint8_t((in[cw - 1] * 3 + in[cw - 2] + 2) >> 2); out[(cw - 1) * 2 + 1] = in[cw - 1];
                    }
                } else if (hs == 2 && vs == 2) {
                    const int near_y = int(y / 2); int far_y = (y & 1) ? near_y + 1 : near_y - 1;
                    far_y = std::min(std::max(far_y, 0), chh - 1);
                    const uint8_t* nr = planes[i].data() + size_t(near_y) * stride; const uint8_t* fr = planes[i].data() + size_t(far_y) * stride;
                    if (cw == 1) { out[0] = out[1] = uint8_t((3 * nr[0] + fr[0] + 2) >> 2); }
                    else {
                        int t0 = 3 * nr[0] + fr[0], t1 = 3 * nr[1] + fr[1];
                        out[0] = uint8_t((t0 * 4 + 8) >> 4); out[1] = uint8_t((t0 * 3 + t1 + 8) >> 4);
                        for (int x = 1; x < cw - 1; ++x) {
                            const int tp = 3 * nr[x - 1] + fr[x - 1], tc = 3 * nr[x] + fr[x], tn = 3 * nr[x + 1] + fr[x + 1];
                            out[2 * x] = uint8_t((3 * tc + tp + 8) >> 4); out[2 * x + 1] = uint8_t((3 * tc + tn + 8) >> 4);
                        }
                        t0 = 3 * nr[cw - 2] + fr[cw - 2]; t1 = 3 * nr[cw - 1] + fr[cw - 1];
                        out[(cw - 1) * 2] = uint8_t((3 * t1 + t0 + 8) >> 4); out[(cw - 1) * 2 + 1] = uint8_t((t1 * 4 + 8) >> 4);
                    }
                } else {   // other integer ratios: sample replication
                    const uint8_t* in = planes[i].data() + size_t(y / vs) * stride;
                    for (uint32_t x = 0; x < W; ++x) out[x] = in[x / hs];
                }
            }
            const bool rgb_direct = adobe_transform == 0 || (comps[0].id == 'R' && comps[1].id == 'G' && comps[2].id == 'B');
            for (uint32_t x = 0; x < W; ++x) {
                uint8_t* o = &rgba[(size_t(y) * W + x) * 4];
                if (rgb_direct) { o[0] = lines[0][x]; o[1] = lines[1][x]; o[2] = lines[2][x]; continue; }
                const float Y = float(lines[0][x]), cb = float(lines[1][x]) - 128.0f, cr = float(lines[2][x]) - 128.0f;
                const float r = Y + 1.40200f * cr, g = Y - 0.34414f * cb - 0.71414f * cr, b = Y + 1.77200f * cb;
                o[0] = clamp_u8(int(r + 0.5f)); o[1] = clamp_u8(int(g + 0.5f)); o[2] = clamp_u8(int(b + 0.5f));
            }
        }
        return true;
    }
};

// ------------------------------------------------------------------------------------------------ imageops::resize(Lanczos3)
inline float sinc_f(float t) {
    if (t == 0.0f) return 1.0f;
    const float a = t * 3.14159274101257324f;   // f32::consts::PI
    return float(std::sin(double(a))) / a;      // sin rounded once to f32 (what a correctly rounded sinf returns)
}
inline float lanczos3(float x) { return std::fabs(x) < 3.0f ? sinc_f(x) * sinc_f(x / 3.0f) : 0.0f; }

struct Taps { uint32_t left; std::vector<float> w; };
Taps taps_for(uint32_t out_i, uint32_t in_n, uint32_t out_n) {
    const float ratio = float(in_n) / float(out_n);
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float support = 3.0f * sratio;
    const float centre = (float(out_i) + 0.5f) * ratio;
    int64_t left = int64_t(std::floor(centre - support)); left = std::min<int64_t>(std::max<int64_t>(left, 0), int64_t(in_n) - 1);
    int64_t right = int64_t(std::ceil(centre + support)); right = std::min<int64_t>(std::max<int64_t>(right, left + 1), int64_t(in_n));
    const float c = centre - 0.5f;
    Taps t; t.left = uint32_t(left);
    float sum = 0.0f;
    for (int64_t i = left; i < right; ++i) { const float w = lanczos3((float(i) - c) / sratio); t.w.push_back(w); sum += w; }
    for (float& w : t.w) w /= sum;
    return t;
}
void resize_lanczos3(const uint8_t* src, uint32_t sw, uint32_t sh, uint32_t dw, uint32_t dh, std::vector<uint8_t>& dst) {
    std::vector<float> tmp(size_t(sw) * dh * 4);   // vertical pass, unrounded
    for (uint32_t oy = 0; oy < dh; ++oy) {
        const Taps t = taps_for(oy, sh, dh);
        for (uint32_t x = 0; x < sw; ++x) {
            float acc[4] = {0, 0, 0, 0};
            for (size_t k = 0; k < t.w.size(); ++k) { const uint8_t* s = src + (size_t(t.left + k) * sw + x) * 4; for (int c = 0; c < 4; ++c) acc[c] += float(s[c]) * t.w[k]; }
            memcpy(&tmp[(size_t(oy) * sw + x) * 4], acc, sizeof acc);
        }
    }
    dst.assign(size_t(dw) * dh * 4, 0);
    for (uint32_t ox = 0; ox < dw; ++ox) {
        const Taps t = taps_for(ox, sw, dw);
        for (uint32_t y = 0; y < dh; ++y) {
            float acc[4] = {0, 0, 0, 0};
            for (size_t k = 0; k < t.w.size(); ++k) { const float* s = &tmp[(size_t(y) * sw + t.left + k) * 4]; for (int c = 0; c < 4; ++c) acc[c] += s[c] * t.w[k]; }
            uint8_t* o = &dst[(size_t(y) * dw + ox) * 4];
            for (int c = 0; c < 4; ++c) { const float v = acc[c] < 0.0f ? 0.0f : acc[c] > 255.0f ? 255.0f : acc[c]; o[c] = uint8_t(std::round(v)); }   // FloatNearest = f32::round
        }
    }
}
}  // namespace


// ------------------------------------------------------------------------------------------------ DDS (image.rs:70-84,285-335 process_dds)
// Upstream hands block-compressed DDS mips to the GPU as they are; this build's textures are RGBA8, so the blocks are decoded here with the
// D3D11 functional-spec arithmetic.  Accepted, like upstream: DX10-header files in BC1_UNORM_SRGB, BC3_UNORM[_SRGB], BC5_UNORM / BC5_SNORM.
namespace {
inline void rgb565(uint32_t c, int out[3]) { const int r = (c >> 11) & 31, g = (c >> 5) & 63, b = c & 31; out[0] = (r << 3) | (r >> 2); out[1] = (g << 2) | (g >> 4); out[2] = (b << 3) | (b >> 2); }
void decode_bc1_colors(const uint8_t* blk, bool opaque_only, uint8_t out[16][4]) {   // 8 bytes: two RGB565 endpoints, 16 x 2-bit indices
    const uint32_t c0 = blk[0] | (blk[1] << 8), c1 = blk[2] | (blk[3] << 8);
    int e0[3], e1[3]; rgb565(c0, e0); rgb565(c1, e1);
    int pal[4][4];
    for (int k = 0; k < 3; ++k) { pal[0][k] = e0[k]; pal[1][k] = e1[k]; }
    pal[0][3] = pal[1][3] = pal[2][3] = pal[3][3] = 255;
    if (c0 > c1 || opaque_only) for (int k = 0; k < 3; ++k) { pal[2][k] = (2 * e0[k] + e1[k] + 1) / 3; pal[3][k] = (e0[k] + 2 * e1[k] + 1) / 3; }
    else { for (int k = 0; k < 3; ++k) { pal[2][k] = (e0[k] + e1[k]) / 2; pal[3][k] = 0; } pal[3][3] = 0; }
    const uint32_t idx = blk[4] | (blk[5] << 8) | (blk[6] << 16) | (uint32_t(blk[7]) << 24);
    for (int t = 0; t < 16; ++t) { const int* c = pal[(idx >> (2 * t)) & 3]; for (int k = 0; k < 4; ++k) out[t][k] = uint8_t(c[k]); }
}
void decode_bc4(const uint8_t* blk, bool snorm, uint8_t out[16]) {   // 8 bytes: two endpoints, 16 x 3-bit indices; SNORM results are mapped to UNORM8 texels
    float a[8];
    if (snorm) { const int s0 = int8_t(blk[0]), s1 = int8_t(blk[1]); a[0] = std::max(s0 / 127.0f, -1.0f); a[1] = std::max(s1 / 127.0f, -1.0f);
        if (s0 > s1) for (int i = 1; i < 7; ++i) a[i + 1] = ((7 - i) * a[0] + i * a[1]) / 7.0f; else { for (int i = 1; i < 5; ++i) a[i + 1] = ((5 - i) * a[0] + i * a[1]) / 5.0f; a[6] = -1.0f; a[7] = 1.0f; } }
    else { const int u0 = blk[0], u1 = blk[1]; a[0] = u0 / 255.0f; a[1] = u1 / 255.0f;
        if (u0 > u1) for (int i = 1; i < 7; ++i) a[i + 1] = ((7 - i) * a[0] + i * a[1]) / 7.0f; else { for (int i = 1; i < 5; ++i) a[i + 1] = ((5 - i) * a[0] + i * a[1]) / 5.0f; a[6] = 0.0f; a[7] = 1.0f; } }
    uint64_t bits = 0; for (int i = 0; i < 6; ++i) bits |= uint64_t(blk[2 + i]) << (8 * i);
    for (int t = 0; t < 16; ++t) { const float v = a[(bits >> (3 * t)) & 7]; const float u = snorm ? v * 0.5f + 0.5f : v; out[t] = uint8_t(std::floor(std::min(std::max(u, 0.0f), 1.0f) * 255.0f + 0.5f)); }
}
}  // namespace

bool decode_dds(const uint8_t* p, size_t n, std::vector<uint8_t>& texels, uint32_t& W, uint32_t& H, uint32_t& mips, uint32_t& srgb) {
    auto u32 = [&](size_t off) { return uint32_t(p[off]) | (uint32_t(p[off + 1]) << 8) | (uint32_t(p[off + 2]) << 16) | (uint32_t(p[off + 3]) << 24); };
    if (n < 128 || memcmp(p, "DDS ", 4) != 0 || u32(4) != 124) { set_error("dds: bad header"); return false; }
    H = u32(12); W = u32(16); mips = std::max(1u, u32(28));
    const uint32_t pf_flags = u32(80), fourcc = u32(84);
    if (!(pf_flags & 4u) || fourcc != 0x30315844u /* "DX10" */ || n < 148) { set_error("dds: only DX10-header files are accepted (as upstream: it matches on the DXGI format)"); return false; }
    const uint32_t dxgi = u32(128);
    enum { BC1, BC3, BC5U, BC5S } kind;
    switch (dxgi) {
        case 72: kind = BC1; srgb = 1; break;      // BC1_UNORM_SRGB
        case 77: kind = BC3; srgb = 0; break;      // BC3_UNORM
        case 78: kind = BC3; srgb = 1; break;      // BC3_UNORM_SRGB
        case 83: kind = BC5U; srgb = 0; break;     // BC5_UNORM
        case 84: kind = BC5S; srgb = 0; break;     // BC5_SNORM
        default: { char m[96]; snprintf(m, sizeof m, "dds: DXGI format %u not supported yet", dxgi); set_error(m); return false; }
    }
    if (!W || !H || W > 16384 || H > 16384 || mips > 15) { set_error("dds: implausible dimensions"); return false; }
    const size_t block_bytes = kind == BC1 ? 8 : 16;
    size_t off = 148; texels.clear();
    for (uint32_t l = 0; l < mips; ++l) {
        const uint32_t w = std::max(1u, W >> l), h = std::max(1u, H >> l);
        const uint32_t bw = (std::max(w, 4u) + 3) / 4, bh = (std::max(h, 4u) + 3) / 4;     // process_dds: (dim >> mip).max(pitch_height)
        if (off + size_t(bw) * bh * block_bytes > n) { set_error("dds: mip data reaches past the end of the file"); return false; }
        const size_t base = texels.size(); texels.resize(base + size_t(w) * h * 4);
        for (uint32_t by = 0; by < bh; ++by) for (uint32_t bx = 0; bx < bw; ++bx) {
            const uint8_t* blk = p + off + (size_t(by) * bw + bx) * block_bytes;
            uint8_t px[16][4];
            if (kind == BC1) decode_bc1_colors(blk, false, px);
            else if (kind == BC3) { uint8_t al[16]; decode_bc4(blk, false, al); decode_bc1_colors(blk + 8, true, px); for (int t = 0; t < 16; ++t) px[t][3] = al[t]; }
            else { uint8_t r[16], g[16]; decode_bc4(blk, kind == BC5S, r); decode_bc4(blk + 8, kind == BC5S, g); for (int t = 0; t < 16; ++t) { px[t][0] = r[t]; px[t][1] = g[t]; px[t][2] = 0; px[t][3] = 255; } }
            for (int t = 0; t < 16; ++t) { const uint32_t x = bx * 4 + (t & 3), y = by * 4 + (t >> 2); if (x < w && y < h) memcpy(&texels[base + (size_t(y) * w + x) * 4], px[t], 4); }
        }
        off += size_t(bw) * bh * block_bytes;
    }
    return true;
}

bool decode_image(const uint8_t* bytes, size_t n, std::vector<uint8_t>& rgba, uint32_t& w, uint32_t& h) {
    static const uint8_t PNG_SIG[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (n >= 8 && !memcmp(bytes, PNG_SIG, 8)) return decode_png(bytes, n, rgba, w, h);
    if (n >= 3 && bytes[0] == 0xff && bytes[1] == 0xd8 && bytes[2] == 0xff) { JpegDecoder d; d.p = bytes; d.n = n; return d.run(rgba, w, h); }
    if (n >= 4 && !memcmp(bytes, "DDS ", 4)) {   // top level of the chain
        std::vector<uint8_t> chain; uint32_t mips, srgb;
        if (!decode_dds(bytes, n, chain, w, h, mips, srgb)) return false;
        rgba.assign(chain.begin(), chain.begin() + size_t(w) * h * 4);
        return true;
    }
    set_error("image: unrecognised container (PNG, JPEG and DX10 DDS are supported)");
    return false;
}

// CreateGpuImage::process_rgba8, TexCompressionMode::None (image.rs:130-283)
void build_mips(const uint8_t* rgba8, uint32_t width, uint32_t height, bool use_mips, const uint32_t* swizzle, std::vector<uint8_t>& out, uint32_t& ow, uint32_t& oh, uint32_t& levels) {
    const uint32_t MAX_SIZE = 2048;
    std::vector<uint8_t> image(rgba8, rgba8 + size_t(width) * height * 4);
    uint32_t w = width, h = height;
    if (w > MAX_SIZE || h > MAX_SIZE) {
        std::vector<uint8_t> r; const uint32_t nw = std::min(w, MAX_SIZE), nh = std::min(h, MAX_SIZE);
        resize_lanczos3(image.data(), w, h, nw, nh, r); image.swap(r); w = nw; h = nh;
    }
    ow = w; oh = h;
    auto mip_count_1d = [](uint32_t e) { uint32_t c = 0; while (e) { ++c; e >>= 1; } return c; };
    levels = use_mips ? std::max(mip_count_1d(w), mip_count_1d(h)) : 1;
    out.clear();
    for (uint32_t l = 0; l < levels; ++l) {
        const size_t base = out.size();
        out.insert(out.end(), image.begin(), image.end());
        // image.rs:214-223 assigns the four channels one after another IN PLACE, so a later channel can pick up an already replaced one
        // ([1,2,0,3] yields (g, b, g, a), not (g, b, r, a)); reproduced as is
        if (swizzle) for (size_t i = base; i < out.size(); i += 4) for (int c = 0; c < 4; ++c) out[i + c] = out[i + (swizzle[c] & 3)];
        if (l + 1 < levels) {
            std::vector<uint8_t> r; const uint32_t nw = std::max(1u, w / 2), nh = std::max(1u, h / 2);
            resize_lanczos3(image.data(), w, h, nw, nh, r); image.swap(r); w = nw; h = nh;
        }
    }
}

}  // namespace kjb_asset_detail
