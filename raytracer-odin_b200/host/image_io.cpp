// image_io.cpp — see image_io.hpp.
#include "image_io.hpp"

#include <zlib.h>

#include <cmath>
#include <cstdio>
#include <cstring>

namespace orh {

bool read_file(const std::string& path, std::vector<uint8_t>* out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (n < 0) { std::fclose(f); return false; }
    out->resize((size_t)n);
    const size_t got = n ? std::fread(out->data(), 1, (size_t)n, f) : 0;
    std::fclose(f);
    return got == (size_t)n;
}

// ------------------------------------------------------------------------------------------------
// PNG
// ------------------------------------------------------------------------------------------------
namespace {

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline int paeth(int a, int b, int c) {
    const int p = a + b - c;
    const int pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    if (pa <= pb && pa <= pc) return a;
    if (pb <= pc) return b;
    return c;
}

} // namespace

bool decode_png(const uint8_t* data, size_t size, Texture* out, std::string* err) {
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (size < 8 || std::memcmp(data, sig, 8) != 0) { *err = "Failed to parse texture"; return false; }
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat, plte, trns;
    bool have_ihdr = false, have_trns = false;
    size_t pos = 8;
    while (pos + 12 <= size) {
        const uint32_t len = be32(data + pos);
        const uint8_t* type = data + pos + 4;
        const uint8_t* body = data + pos + 8;
        if (pos + 12 + (size_t)len > size) { *err = "Failed to parse texture (truncated PNG chunk)"; return false; }
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len != 13) { *err = "Failed to parse texture (bad IHDR)"; return false; }
            w = be32(body); h = be32(body + 4);
            depth = body[8]; ctype = body[9]; interlace = body[12];
            have_ihdr = true;
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(body, body + len);
        } else if (!std::memcmp(type, "tRNS", 4)) {
            trns.assign(body, body + len);
            have_trns = true;
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), body, body + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || w == 0 || h == 0 || w > (1u << 24) || h > (1u << 24)) { *err = "Failed to parse texture (bad PNG header)"; return false; }
    if (interlace != 0) { *err = "Failed to parse texture (interlaced PNG is not supported)"; return false; }
    int src_ch;
    switch (ctype) {
    case 0: src_ch = 1; break;
    case 2: src_ch = 3; break;
    case 3: src_ch = 1; break;
    case 4: src_ch = 2; break;
    case 6: src_ch = 4; break;
    default: *err = "Failed to parse texture (bad PNG colour type)"; return false;
    }
    if (!(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16) || (depth < 8 && ctype != 0 && ctype != 3) ||
        (depth == 16 && ctype == 3)) {
        *err = "Failed to parse texture (bad PNG bit depth)"; return false;
    }
    const size_t bits_pp = (size_t)src_ch * (size_t)depth;
    const size_t row_bytes = ((size_t)w * bits_pp + 7) / 8;
    const size_t bpp = bits_pp >= 8 ? bits_pp / 8 : 1;
    std::vector<uint8_t> raw((row_bytes + 1) * (size_t)h);
    {
        uLongf dst_len = (uLongf)raw.size();
        const int rc = uncompress(raw.data(), &dst_len, idat.data(), (uLong)idat.size());
        if (rc != Z_OK || dst_len != raw.size()) { *err = "Failed to parse texture (PNG inflate)"; return false; }
    }
    // undo the scanline filters in place
    std::vector<uint8_t> zero(row_bytes, 0);
    for (size_t y = 0; y < h; y++) {
        uint8_t* row = raw.data() + y * (row_bytes + 1) + 1;
        const uint8_t* up = y ? raw.data() + (y - 1) * (row_bytes + 1) + 1 : zero.data();
        const int ft = row[-1];
        for (size_t x = 0; x < row_bytes; x++) {
            const int a = x >= bpp ? row[x - bpp] : 0, b = up[x], c = x >= bpp ? up[x - bpp] : 0;
            int v = row[x];
            switch (ft) {
            case 0: break;
            case 1: v += a; break;
            case 2: v += b; break;
            case 3: v += (a + b) >> 1; break;
            case 4: v += paeth(a, b, c); break;
            default: *err = "Failed to parse texture (bad PNG filter)"; return false;
            }
            row[x] = (uint8_t)v;
        }
    }
    // native channel count, like stbi_load(..., req_comp = 0)
    int out_ch = ctype == 3 ? (have_trns ? 4 : 3) : src_ch + ((have_trns && (ctype == 0 || ctype == 2)) ? 1 : 0);
    out->is_f32 = false;
    out->width = (int)w; out->height = (int)h; out->channels = out_ch;
    out->f32.clear();
    out->u8.assign((size_t)w * h * (size_t)out_ch, 255);
    static const int depth_scale[9] = {0, 0xff, 0x55, 0, 0x11, 0, 0, 0, 0x01};
    for (size_t y = 0; y < h; y++) {
        const uint8_t* row = raw.data() + y * (row_bytes + 1) + 1;
        uint8_t* dst = out->u8.data() + y * (size_t)w * (size_t)out_ch;
        for (size_t x = 0; x < w; x++) {
            uint16_t s[4] = {0, 0, 0, 0}; // source samples at file precision
            if (depth == 16) {
                for (int c = 0; c < src_ch; c++) s[c] = (uint16_t)((row[(x * src_ch + c) * 2] << 8) | row[(x * src_ch + c) * 2 + 1]);
            } else if (depth == 8) {
                for (int c = 0; c < src_ch; c++) s[c] = row[x * src_ch + c];
            } else {
                const size_t bit = x * (size_t)depth;
                s[0] = (uint16_t)((row[bit >> 3] >> (8 - depth - (int)(bit & 7))) & ((1 << depth) - 1));
            }
            if (ctype == 3) {
                const size_t idx = s[0];
                for (int c = 0; c < 3; c++) dst[x * out_ch + c] = idx * 3 + c < plte.size() ? plte[idx * 3 + c] : 0;
                if (out_ch == 4) dst[x * 4 + 3] = idx < trns.size() ? trns[idx] : 255;
                continue;
            }
            for (int c = 0; c < src_ch; c++) {
                uint8_t v;
                if (depth == 16) v = (uint8_t)(s[c] >> 8); // stbi__convert_16_to_8 keeps the high byte
                else if (depth == 8) v = (uint8_t)s[c];
                else v = (uint8_t)(s[c] * depth_scale[depth]);
                dst[x * out_ch + c] = v;
            }
            if (out_ch == src_ch + 1) { // colour-key transparency
                bool key = true;
                for (int c = 0; c < src_ch && key; c++) {
                    const uint16_t k = (size_t)(2 * c + 1) < trns.size() ? (uint16_t)((trns[2 * c] << 8) | trns[2 * c + 1]) : 0;
                    key = (depth == 16 ? s[c] : (uint16_t)(s[c] & 0xffff)) == k;
                }
                dst[x * out_ch + src_ch] = key ? 0 : 255;
            }
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// Radiance RGBE
// ------------------------------------------------------------------------------------------------
namespace {

bool hdr_line(const uint8_t* data, size_t size, size_t* pos, std::string* line) {
    line->clear();
    if (*pos >= size) return false;
    while (*pos < size && data[*pos] != '\n') { line->push_back((char)data[*pos]); (*pos)++; }
    if (*pos < size) (*pos)++;
    return true;
}

inline void rgbe_to_float(const uint8_t* in, float* o) { // stbi__hdr_convert, req_comp = 3
    if (in[3] != 0) {
        const float f1 = std::ldexp(1.0f, (int)in[3] - (128 + 8));
        o[0] = in[0] * f1; o[1] = in[1] * f1; o[2] = in[2] * f1;
    } else {
        o[0] = o[1] = o[2] = 0.0f;
    }
}

} // namespace

bool decode_hdr(const uint8_t* data, size_t size, Texture* out, std::string* err) {
    size_t pos = 0;
    std::string line;
    if (!hdr_line(data, size, &pos, &line) || (line != "#?RADIANCE" && line != "#?RGBE")) { *err = "Failed to parse texture"; return false; }
    bool valid = false;
    for (;;) {
        if (!hdr_line(data, size, &pos, &line)) { *err = "Failed to parse texture (HDR header)"; return false; }
        if (line.empty()) break;
        if (line == "FORMAT=32-bit_rle_rgbe") valid = true;
    }
    if (!valid) { *err = "Failed to parse texture (unsupported HDR format)"; return false; }
    if (!hdr_line(data, size, &pos, &line)) { *err = "Failed to parse texture (HDR resolution)"; return false; }
    int h = 0, w = 0;
    if (std::sscanf(line.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) {
        *err = "Failed to parse texture (unsupported HDR data layout)"; return false;
    }
    out->is_f32 = true;
    out->width = w; out->height = h; out->channels = 3;
    out->u8.clear();
    out->f32.assign((size_t)w * h * 3, 0.0f);
    auto need = [&](size_t n) { return pos + n <= size; };
    auto flat_from = [&](size_t first_pixel) -> bool { // flat RGBE pixels from `first_pixel` on
        for (size_t i = first_pixel; i < (size_t)w * h; i++) {
            if (!need(4)) return false;
            rgbe_to_float(data + pos, &out->f32[i * 3]);
            pos += 4;
        }
        return true;
    };
    if (w < 8 || w >= 32768) {
        if (!flat_from(0)) { *err = "Failed to parse texture (HDR data)"; return false; }
        return true;
    }
    std::vector<uint8_t> scan((size_t)w * 4);
    for (int j = 0; j < h; j++) {
        if (!need(4)) { *err = "Failed to parse texture (HDR data)"; return false; }
        const int c1 = data[pos], c2 = data[pos + 1], len = data[pos + 2];
        if (c1 != 2 || c2 != 2 || (len & 0x80)) {
            // not run-length encoded: stb only accepts this for the whole image, from the first pixel
            if (j != 0) { *err = "Failed to parse texture (HDR mixes flat and RLE data)"; return false; }
            if (!flat_from(0)) { *err = "Failed to parse texture (HDR data)"; return false; }
            return true;
        }
        const int width_in = (len << 8) | data[pos + 3];
        pos += 4;
        if (width_in != w) { *err = "Failed to parse texture (HDR scanline width)"; return false; }
        for (int k = 0; k < 4; k++) {
            int i = 0;
            while (i < w) {
                if (!need(1)) { *err = "Failed to parse texture (HDR data)"; return false; }
                int count = data[pos++];
                if (count > 128) {
                    count -= 128;
                    if (count == 0 || count > w - i || !need(1)) { *err = "Failed to parse texture (HDR run)"; return false; }
                    const uint8_t v = data[pos++];
                    for (int z = 0; z < count; z++) scan[(size_t)(i++) * 4 + k] = v;
                } else {
                    if (count == 0 || count > w - i || !need((size_t)count)) { *err = "Failed to parse texture (HDR run)"; return false; }
                    for (int z = 0; z < count; z++) scan[(size_t)(i++) * 4 + k] = data[pos++];
                }
            }
        }
        for (int i = 0; i < w; i++) rgbe_to_float(&scan[(size_t)i * 4], &out->f32[((size_t)j * w + i) * 3]);
    }
    return true;
}

bool load_texture(const std::string& path, Texture* out, std::string* err) {
    std::vector<uint8_t> bytes;
    if (!read_file(path, &bytes)) { *err = "Failed to read texture file: " + path; return false; } // textures.odin:28
    const bool is_hdr = bytes.size() >= 6 && (!std::memcmp(bytes.data(), "#?RADIANCE", std::min<size_t>(10, bytes.size())) ||
                                              !std::memcmp(bytes.data(), "#?RGBE", 6));
    if (is_hdr) return decode_hdr(bytes.data(), bytes.size(), out, err);
    return decode_png(bytes.data(), bytes.size(), out, err);
}

// ------------------------------------------------------------------------------------------------
// writers
// ------------------------------------------------------------------------------------------------
bool write_ppm(const std::string& path, int w, int h, const uint8_t* rgb, std::string* err) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) { *err = "Failed to open file " + path; return false; }
    std::fprintf(f, "P6\n%d %d\n255\n", w, h); // output.odin:92
    const size_t n = (size_t)w * h * 3;
    const bool ok = std::fwrite(rgb, 1, n, f) == n;
    std::fclose(f);
    if (!ok) *err = "short write to " + path;
    return ok;
}

bool write_png(const std::string& path, int w, int h, const uint8_t* rgb, std::string* err) {
    const size_t row = (size_t)w * 3;
    std::vector<uint8_t> raw((row + 1) * (size_t)h);
    for (int y = 0; y < h; y++) {
        raw[(size_t)y * (row + 1)] = 0; // filter: none
        std::memcpy(&raw[(size_t)y * (row + 1) + 1], rgb + (size_t)y * row, row);
    }
    uLongf clen = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(clen);
    if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) { *err = "PNG deflate failed"; return false; }
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) { *err = "Failed to open file " + path; return false; }
    auto chunk = [&](const char* type, const uint8_t* body, uint32_t len) {
        uint8_t hdr[8] = {(uint8_t)(len >> 24), (uint8_t)(len >> 16), (uint8_t)(len >> 8), (uint8_t)len,
                          (uint8_t)type[0], (uint8_t)type[1], (uint8_t)type[2], (uint8_t)type[3]};
        std::fwrite(hdr, 1, 8, f);
        if (len) std::fwrite(body, 1, len, f);
        uLong crc = crc32(0L, hdr + 4, 4);
        if (len) crc = crc32(crc, body, len);
        const uint8_t c[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
        std::fwrite(c, 1, 4, f);
    };
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::fwrite(sig, 1, 8, f);
    uint8_t ihdr[13] = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w,
                        (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h, 8, 2, 0, 0, 0};
    chunk("IHDR", ihdr, 13);
    chunk("IDAT", comp.data(), (uint32_t)clen);
    chunk("IEND", nullptr, 0);
    const bool ok = std::fclose(f) == 0;
    if (!ok) *err = "short write to " + path;
    return ok;
}

} // namespace orh
