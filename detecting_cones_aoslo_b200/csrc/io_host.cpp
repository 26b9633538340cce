// io_host.cpp -- host-side input path of the detection run: the image / ground-truth readers the reference
// reaches through cv2.imread(path, -1) and pd.read_csv(path).to_numpy() (pipeline_with_delitions.py:443-446,
// main_with_delitions.py:30-33, :153-155).  Decodes straight into caller-provided (pinned) buffers so that a batch
// of frames goes to the device with one asynchronous copy.
//
// TIFF subset: baseline, 8-bit, one sample per pixel, BlackIsZero, strips, uncompressed or PackBits (the shipped
// AOSLO crop is PackBits, 15 rows per strip), either byte order.  Anything else -> AOSLO_ERR_UNSUPPORTED.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "aoslo_b200.h"

namespace {

struct Reader {
    std::vector<uint8_t> buf;
    bool le = true;
    bool load(const char* path) {
        FILE* f = fopen(path, "rb");
        if (!f) return false;
        fseek(f, 0, SEEK_END);
        long n = ftell(f);
        fseek(f, 0, SEEK_SET);
        if (n < 8) { fclose(f); return false; }
        buf.resize((size_t)n);
        size_t got = fread(buf.data(), 1, (size_t)n, f);
        fclose(f);
        return got == (size_t)n;
    }
    bool ok(size_t off, size_t len) const { return off <= buf.size() && len <= buf.size() - off; }
    uint16_t u16(size_t o) const {
        return le ? (uint16_t)(buf[o] | (buf[o + 1] << 8)) : (uint16_t)((buf[o] << 8) | buf[o + 1]);
    }
    uint32_t u32(size_t o) const {
        return le ? ((uint32_t)buf[o] | ((uint32_t)buf[o + 1] << 8) | ((uint32_t)buf[o + 2] << 16) | ((uint32_t)buf[o + 3] << 24))
                  : (((uint32_t)buf[o] << 24) | ((uint32_t)buf[o + 1] << 16) | ((uint32_t)buf[o + 2] << 8) | (uint32_t)buf[o + 3]);
    }
};

struct TiffDir {
    uint32_t width = 0, height = 0, bits = 1, compression = 1, photometric = 1, samples = 1, rows_per_strip = 0xffffffffu,
             planar = 1;
    std::vector<uint32_t> strip_off, strip_len;
    bool tiled = false;
};

// values of an IFD entry of type SHORT (3) or LONG (4)
bool entry_values(const Reader& r, size_t e, std::vector<uint32_t>& out) {
    const uint16_t type = r.u16(e + 2);
    const uint32_t count = r.u32(e + 4);
    const size_t sz = type == 3 ? 2 : (type == 4 ? 4 : 0);
    if (!sz || count == 0 || count > (1u << 24)) return false;
    size_t at = e + 8;
    if ((size_t)count * sz > 4) {
        at = r.u32(e + 8);
        if (!r.ok(at, (size_t)count * sz)) return false;
    }
    out.resize(count);
    for (uint32_t i = 0; i < count; ++i) out[i] = sz == 2 ? r.u16(at + 2 * (size_t)i) : r.u32(at + 4 * (size_t)i);
    return true;
}

int parse(const Reader& rd, TiffDir& d) {
    Reader& r = const_cast<Reader&>(rd);
    if (r.buf[0] == 'I' && r.buf[1] == 'I') r.le = true;
    else if (r.buf[0] == 'M' && r.buf[1] == 'M') r.le = false;
    else return AOSLO_ERR_INVALID;
    if (r.u16(2) != 42) return AOSLO_ERR_UNSUPPORTED;          // BigTIFF etc.
    const size_t ifd = r.u32(4);
    if (!r.ok(ifd, 2)) return AOSLO_ERR_INVALID;
    const uint16_t n = r.u16(ifd);
    if (!r.ok(ifd + 2, (size_t)n * 12)) return AOSLO_ERR_INVALID;
    for (uint16_t i = 0; i < n; ++i) {
        const size_t e = ifd + 2 + 12 * (size_t)i;
        const uint16_t tag = r.u16(e);
        std::vector<uint32_t> v;
        switch (tag) {
            case 256: if (entry_values(r, e, v)) d.width = v[0]; break;
            case 257: if (entry_values(r, e, v)) d.height = v[0]; break;
            case 258: if (entry_values(r, e, v)) d.bits = v[0]; break;
            case 259: if (entry_values(r, e, v)) d.compression = v[0]; break;
            case 262: if (entry_values(r, e, v)) d.photometric = v[0]; break;
            case 273: if (!entry_values(r, e, d.strip_off)) return AOSLO_ERR_INVALID; break;
            case 277: if (entry_values(r, e, v)) d.samples = v[0]; break;
            case 278: if (entry_values(r, e, v)) d.rows_per_strip = v[0]; break;
            case 279: if (!entry_values(r, e, d.strip_len)) return AOSLO_ERR_INVALID; break;
            case 284: if (entry_values(r, e, v)) d.planar = v[0]; break;
            case 322: case 323: case 324: case 325: d.tiled = true; break;
            default: break;
        }
    }
    if (!d.width || !d.height || d.width > 65535u || d.height > 65535u) return AOSLO_ERR_INVALID;
    return AOSLO_OK;
}

}  // namespace

extern "C" int aoslo_tiff_info(const char* path, int* h_H, int* h_W, int* h_bits, int* h_samples) {
    if (!path || !h_H || !h_W) return AOSLO_ERR_INVALID;
    Reader r;
    if (!r.load(path)) return AOSLO_ERR_INVALID;
    TiffDir d;
    int st = parse(r, d);
    if (st != AOSLO_OK) return st;
    *h_H = (int)d.height;
    *h_W = (int)d.width;
    if (h_bits) *h_bits = (int)d.bits;
    if (h_samples) *h_samples = (int)d.samples;
    return AOSLO_OK;
}

extern "C" int aoslo_tiff_read_gray8(const char* path, uint8_t* h_out, int out_pitch_bytes) {
    if (!path || !h_out) return AOSLO_ERR_INVALID;
    Reader r;
    if (!r.load(path)) return AOSLO_ERR_INVALID;
    TiffDir d;
    int st = parse(r, d);
    if (st != AOSLO_OK) return st;
    if (d.bits != 8 || d.samples != 1 || d.photometric != 1 || d.tiled || d.planar != 1) return AOSLO_ERR_UNSUPPORTED;
    if (d.compression != 1 && d.compression != 32773) return AOSLO_ERR_UNSUPPORTED;
    if (out_pitch_bytes < (int)d.width) return AOSLO_ERR_INVALID;
    if (d.strip_off.empty() || d.strip_off.size() != d.strip_len.size()) return AOSLO_ERR_INVALID;
    const uint32_t rps = d.rows_per_strip < d.height ? d.rows_per_strip : d.height;
    if (rps == 0 || (size_t)((d.height + rps - 1) / rps) != d.strip_off.size()) return AOSLO_ERR_INVALID;
    std::vector<uint8_t> strip;
    for (size_t s = 0; s < d.strip_off.size(); ++s) {
        const uint32_t row0 = (uint32_t)s * rps;
        const uint32_t rows = (row0 + rps <= d.height) ? rps : d.height - row0;
        const size_t want = (size_t)rows * d.width;
        if (!r.ok(d.strip_off[s], d.strip_len[s])) return AOSLO_ERR_INVALID;
        const uint8_t* src = r.buf.data() + d.strip_off[s];
        const uint8_t* px;
        if (d.compression == 1) {
            if (d.strip_len[s] < want) return AOSLO_ERR_INVALID;
            px = src;
        } else {                                   // PackBits (TIFF 6.0 section 9)
            strip.assign(want, 0);
            size_t o = 0, i = 0;
            const size_t n = d.strip_len[s];
            while (i < n && o < want) {
                const int8_t c = (int8_t)src[i++];
                if (c >= 0) {
                    size_t len = (size_t)c + 1;
                    if (i + len > n) return AOSLO_ERR_INVALID;
                    if (o + len > want) len = want - o;
                    memcpy(strip.data() + o, src + i, len);
                    i += (size_t)c + 1;
                    o += len;
                } else if (c != -128) {
                    size_t len = (size_t)(1 - (int)c);
                    if (i >= n) return AOSLO_ERR_INVALID;
                    if (o + len > want) len = want - o;
                    memset(strip.data() + o, src[i++], len);
                    o += len;
                }
            }
            if (o != want) return AOSLO_ERR_INVALID;
            px = strip.data();
        }
        for (uint32_t y = 0; y < rows; ++y)
            memcpy(h_out + (size_t)(row0 + y) * out_pitch_bytes, px + (size_t)y * d.width, d.width);
    }
    return AOSLO_OK;
}

// "x,y" per line (float text).  skip_first_row = 1 reproduces pd.read_csv's default, which consumes the first line of
// the header-less annotation file as column names (the reference therefore uses 7 627 of the 7 628 points).
extern "C" int aoslo_csv_read_points(const char* path, int skip_first_row, double* h_xy_out, int capacity, int* h_n_out) {
    if (!path || !h_n_out || capacity < 0 || (capacity > 0 && !h_xy_out)) return AOSLO_ERR_INVALID;
    FILE* f = fopen(path, "rb");
    if (!f) return AOSLO_ERR_INVALID;
    std::string text;
    char chunk[1 << 16];
    size_t got;
    while ((got = fread(chunk, 1, sizeof(chunk), f)) > 0) text.append(chunk, got);
    fclose(f);
    int n = 0, line_no = 0;
    size_t pos = 0;
    int status = AOSLO_OK;
    while (pos < text.size()) {
        size_t eol = text.find('\n', pos);
        if (eol == std::string::npos) eol = text.size();
        std::string line = text.substr(pos, eol - pos);
        pos = eol + 1;
        while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
        if (line.empty()) continue;
        if (line_no++ == 0 && skip_first_row) continue;
        char* end = nullptr;
        const char* c = line.c_str();
        const double x = strtod(c, &end);
        if (end == c) { status = AOSLO_ERR_INVALID; break; }
        while (*end == ' ' || *end == '\t') ++end;
        if (*end != ',' && *end != ';') { status = AOSLO_ERR_INVALID; break; }
        const char* c2 = end + 1;
        const double y = strtod(c2, &end);
        if (end == c2) { status = AOSLO_ERR_INVALID; break; }
        if (n < capacity) { h_xy_out[2 * (size_t)n] = x; h_xy_out[2 * (size_t)n + 1] = y; }
        ++n;
    }
    *h_n_out = n;                                   // the number of points in the file, even beyond the capacity
    if (status != AOSLO_OK) return status;
    return n > capacity ? AOSLO_ERR_CAPACITY : AOSLO_OK;
}
